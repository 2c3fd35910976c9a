set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r7
(time timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_headline.py -m gpu -q -x -k "not accuracy and not forced" 2>&1 | tail -30) > gpurun_out/r7/pytest_tc.log 2>&1
(DCLL_WG2_TILE=16 timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -k "weight_gradient or training_step" 2>&1 | tail -30) > gpurun_out/r7/pytest_tc16.log 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > gpurun_out/r7/b_t8.json 2> gpurun_out/r7/b_t8.err
DCLL_WG2_TILE=16 timeout 300 python bench.py $Q > gpurun_out/r7/b_t16.json 2> gpurun_out/r7/b_t16.err
for d in 3 4 7; do
DCLL_WG2_DEBUG=$d timeout 300 python bench.py --timesteps 64 --burnin 4 --steps 2 --warmup 1 --no-cpu --no-extras --profile-every 5 > gpurun_out/r7/b_t8_dbg$d.json 2> gpurun_out/r7/b_t8_dbg$d.err
DCLL_WG2_TILE=16 DCLL_WG2_DEBUG=$d timeout 300 python bench.py --timesteps 64 --burnin 4 --steps 2 --warmup 1 --no-cpu --no-extras --profile-every 5 > gpurun_out/r7/b_t16_dbg$d.json 2> gpurun_out/r7/b_t16_dbg$d.err
done
echo done
