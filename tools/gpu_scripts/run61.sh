set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r61
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_readout.py tests/test_gpu_f16x2.py tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -x -q -k "not opt_in" > $O/pytest.log 2>&1
tail -3 $O/pytest.log
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > $O/b.json 2> $O/b.err
echo done
