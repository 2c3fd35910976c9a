set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r40
mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_readout.py tests/test_gpu_tensorcore.py tests/test_gpu_f16x2.py -m gpu -q -x 2>&1 | tail -6) > $O/pytest.log 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > $O/b_f16.json 2> $O/b_f16.err
echo done
