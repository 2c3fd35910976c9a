set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r13
mkdir -p $O
timeout 300 python tools/timeline.py > $O/tl_default.txt 2>&1
DCLL_CONV_MMA2=3 timeout 300 python tools/timeline.py > $O/tl_mma2_all_fused.txt 2>&1
DCLL_CONV_MMA2=3 DCLL_TRACE_FUSE=0 timeout 300 python tools/timeline.py > $O/tl_mma2_all_nofuse.txt 2>&1
(timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -k "forward_teacher_forced or training_step or window_equals" 2>&1 | tail -5) > $O/pytest_tc.log 2>&1
echo done
