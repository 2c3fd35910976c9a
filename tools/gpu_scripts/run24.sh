set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r24
mkdir -p $O
export DCLL_PRECISION=f16x2
timeout 300 python tools/tc_check.py --wgrad > $O/check_f16.txt 2>&1
DCLL_CONV_MMA2=2 timeout 300 python tools/tc_check.py > $O/check_f16_mma2.txt 2>&1
DCLL_WG2_PAIR=0 timeout 300 python tools/tc_check.py --wgrad > $O/check_f16_nopair.txt 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py --precision f16x2 $Q > $O/b_f16.json 2> $O/b_f16.err
DCLL_CONV_MMA2=3 DCLL_TRACE_FUSE=0 timeout 300 python bench.py --precision f16x2 $Q > $O/b_f16_all2_nofuse.json 2> $O/b_f16_all2_nofuse.err
DCLL_CONV_MMA2=3 timeout 300 python bench.py --precision f16x2 $Q > $O/b_f16_all2_fuse.json 2> $O/b_f16_all2_fuse.err
echo done
