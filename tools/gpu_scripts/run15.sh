set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r15
mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_headline.py -m gpu -q -x -k "not accuracy" 2>&1 | tail -8) > $O/pytest_tc.log 2>&1
(DCLL_CONV_MMA2=2 timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -k "forward_teacher_forced or training_step or window_equals" 2>&1 | tail -8) > $O/pytest_mma2.log 2>&1
(DCLL_TRACE_FUSE=2 timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -k "forward_teacher_forced or training_step or window_equals" 2>&1 | tail -8) > $O/pytest_fuse2.log 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > $O/b_default.json 2> $O/b_default.err
DCLL_CONV_MMA2=3 timeout 300 python bench.py $Q > $O/b_all2_fuse.json 2> $O/b_all2_fuse.err
DCLL_TRACE_FUSE=2 timeout 300 python bench.py $Q > $O/b_fuse2.json 2> $O/b_fuse2.err
DCLL_CONV_MMA2=3 DCLL_TRACE_FUSE=2 timeout 300 python bench.py $Q > $O/b_all2_fuse2.json 2> $O/b_all2_fuse2.err
DCLL_TIMELINE_LAYER=1 timeout 300 python tools/timeline.py > $O/tl_default_l1.txt 2>&1
DCLL_TIMELINE_LAYER=1 DCLL_CONV_MMA2=3 timeout 300 python tools/timeline.py > $O/tl_all2_l1.txt 2>&1
DCLL_TIMELINE_LAYER=0 DCLL_TRACE_FUSE=2 timeout 300 python tools/timeline.py > $O/tl_fuse2_l0.txt 2>&1
echo done
