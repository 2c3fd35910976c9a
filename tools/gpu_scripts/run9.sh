set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r9
(time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/r9/pytest_all.log 2>&1
python __graft_entry__.py --smoke > gpurun_out/r9/smoke.log 2>&1; echo smoke rc=$? >> gpurun_out/r9/smoke.log
(time python bench.py --steps 3 --warmup 3) > gpurun_out/r9/b_full.json 2> gpurun_out/r9/b_full.err
echo done
