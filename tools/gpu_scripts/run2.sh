set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
(time timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -s -k "weight_gradient or training_step or window_equals or other_padding" 2>&1 | tail -60) > gpurun_out/r2/pytest_tc.log 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > gpurun_out/r2/b_tc2.json 2> gpurun_out/r2/b_tc2.err
DCLL_WGRAD_TC2=0 timeout 300 python bench.py $Q > gpurun_out/r2/b_tc1.json 2> gpurun_out/r2/b_tc1.err
(time timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -80) > gpurun_out/r2/pytest_all.log 2>&1
echo done
