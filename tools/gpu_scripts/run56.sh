set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r56
mkdir -p $O
(time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30) > $O/pytest_all.log 2>&1
python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo smoke rc=$? >> $O/smoke.log
(time python bench.py --steps 5 --warmup 3) > $O/b_full.json 2> $O/b_full.err
echo done
