set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r26
mkdir -p $O
Q="--timesteps 64 --burnin 4 --steps 2 --warmup 1 --no-cpu --no-extras --profile-every 5 --precision f16x2"
for d in 0 3 4 7; do
DCLL_WG2_PAIR=0 DCLL_WG2_DEBUG=$d timeout 300 python bench.py $Q > $O/b_np_dbg$d.json 2> $O/b_np_dbg$d.err
done
DCLL_WG2_NA=37 timeout 300 python bench.py $Q > $O/b_na37.json 2> $O/b_na37.err
timeout 300 python bench.py $Q > $O/b_pair.json 2> $O/b_pair.err
for d in 1 2 4 7; do
DCLL_CONV_MMA2=3 DCLL_TRACE_FUSE=0 DCLL_CONV_DEBUG=$d timeout 300 python bench.py $Q > $O/b_conv_dbg$d.json 2> $O/b_conv_dbg$d.err
done
echo done
