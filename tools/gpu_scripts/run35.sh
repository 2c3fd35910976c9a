set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r35
mkdir -p $O
timeout 120 tools/_build/mma_bench > $O/mma_bench.txt 2>&1
(timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_f16x2.py tests/test_gpu_headline.py tests/test_gpu_shapes.py -m gpu -q -x -k "not accuracy" 2>&1 | tail -8) > $O/pytest.log 2>&1
(DCLL_WGRAD_TC2=0 timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -k "weight_gradient or training_step or window_equals" 2>&1 | tail -8) > $O/pytest_wgtc.log 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > $O/b_f16.json 2> $O/b_f16.err
echo done
