set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r60
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_f16x2.py -m gpu -x -q -k "opt_in or window_equals" > $O/pytest.log 2>&1
tail -5 $O/pytest.log
echo done
