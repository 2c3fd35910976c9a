set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r32
mkdir -p $O
(time timeout 1500 python -m pytest tests -m gpu -q -s -k "accuracy" 2>&1 | tail -40) > $O/pytest_acc.log 2>&1
(time timeout 1500 python -m pytest tests -m gpu -q -k "not accuracy" 2>&1 | tail -15) > $O/pytest_all.log 2>&1
(time python bench.py --precision f16x2 --steps 3 --warmup 3 --no-extras) > $O/b_full_f16.json 2> $O/b_full_f16.err
echo done
