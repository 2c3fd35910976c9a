set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r23
mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_readout.py tests/test_gpu_tensorcore.py tests/test_gpu_headline.py -m gpu -q -x -k "not accuracy" 2>&1 | tail -8) > $O/pytest.log 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > $O/b_default.json 2> $O/b_default.err
echo done
