set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r5
(time timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_headline.py -m gpu -q -x -k "not accuracy and not forced" 2>&1 | tail -30) > gpurun_out/r5/pytest_tc.log 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > gpurun_out/r5/b_default.json 2> gpurun_out/r5/b_default.err
DCLL_CONV_MMA2=3 DCLL_CONV_TMA=2 timeout 300 python bench.py $Q > gpurun_out/r5/b_mma2all_fused_tma.json 2> gpurun_out/r5/b_mma2all_fused_tma.err
DCLL_CONV_MMA2=3 timeout 300 python bench.py $Q > gpurun_out/r5/b_mma2all_fused.json 2> gpurun_out/r5/b_mma2all_fused.err
for d in 3 4 7; do
DCLL_WG2_DEBUG=$d timeout 300 python bench.py --timesteps 64 --burnin 4 --steps 2 --warmup 1 --no-cpu --no-extras --profile-every 5 > gpurun_out/r5/b_dbg$d.json 2> gpurun_out/r5/b_dbg$d.err
done
echo done
