set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r18
mkdir -p $O
(timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -k "training_step or window_equals" 2>&1 | tail -15) > $O/pytest_tc.log 2>&1
(timeout 600 python -m pytest tests/test_gpu_headline.py -m gpu -q -x -k "not accuracy" 2>&1 | tail -15) > $O/pytest_head.log 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > $O/b_default.json 2> $O/b_default.err
DCLL_WG2_PAIR=0 timeout 300 python bench.py $Q > $O/b_nopair.json 2> $O/b_nopair.err
timeout 300 python tools/timeline.py --timesteps 32 > $O/tl_default.txt 2>&1
echo done
