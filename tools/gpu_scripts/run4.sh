set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r4
(time timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x 2>&1 | tail -30) > gpurun_out/r4/pytest_tc.log 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > gpurun_out/r4/b_tma.json 2> gpurun_out/r4/b_tma.err
DCLL_WG2_TMA=0 DCLL_CONV_TMA=0 timeout 300 python bench.py $Q > gpurun_out/r4/b_notma.json 2> gpurun_out/r4/b_notma.err
DCLL_CONV_TMA=0 timeout 300 python bench.py $Q > gpurun_out/r4/b_wgtma_only.json 2> gpurun_out/r4/b_wgtma_only.err
for d in 3 4 7; do
DCLL_WG2_DEBUG=$d timeout 300 python bench.py --timesteps 64 --burnin 4 --steps 2 --warmup 1 --no-cpu --no-extras --profile-every 5 > gpurun_out/r4/b_dbg$d.json 2> gpurun_out/r4/b_dbg$d.err
done
(time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/r4/pytest_all.log 2>&1
echo done
