set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r59
mkdir -p $O
DCLL_RB_TC=1 timeout 300 python -m pytest tests/test_gpu_f16x2.py -m gpu -x -q > $O/pytest.log 2>&1
tail -12 $O/pytest.log
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
DCLL_RB_TC=1 timeout 300 python bench.py $Q > $O/b_tc.json 2> $O/b_tc.err
echo done
