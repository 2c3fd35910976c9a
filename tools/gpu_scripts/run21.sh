set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r21
mkdir -p $O
timeout 120 tools/_build/mma_bench > $O/mma_bench.txt 2>&1
(timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_tensorcore.py -m gpu -q -x -k "not accuracy" 2>&1 | tail -8) > $O/pytest_head.log 2>&1
timeout 300 python tools/timeline.py --timesteps 32 > $O/tl_default.txt 2>&1

Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > $O/b_default.json 2> $O/b_default.err
DCLL_WG2_NA=38 timeout 300 python bench.py $Q > $O/b_na38.json 2> $O/b_na38.err
