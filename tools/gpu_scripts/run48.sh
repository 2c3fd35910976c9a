set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r48
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_f16x2.py -m gpu -x -q > $O/pytest.log 2>&1
tail -3 $O/pytest.log
(time python bench.py --steps 3 --warmup 3 --no-cpu --no-extras) > $O/b_full.json 2> $O/b_full.err
echo done
