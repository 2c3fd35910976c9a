set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r11
(DCLL_CONV_MMA2=2 timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -k "forward_teacher_forced or training_step or window_equals" 2>&1 | tail -8) > gpurun_out/r11/pytest_mma2.log 2>&1
(timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_headline.py -m gpu -q -x -k "not accuracy and not forced" 2>&1 | tail -8) > gpurun_out/r11/pytest_tc.log 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > gpurun_out/r11/b_default.json 2> gpurun_out/r11/b_default.err
DCLL_CONV_MMA2=3 DCLL_TRACE_FUSE=0 timeout 300 python bench.py $Q > gpurun_out/r11/b_all2_nofuse.json 2> gpurun_out/r11/b_all2_nofuse.err
DCLL_CONV_MMA2=3 timeout 300 python bench.py $Q > gpurun_out/r11/b_all2_fuse.json 2> gpurun_out/r11/b_all2_fuse.err
DCLL_CONV_DEBUG=3 timeout 300 python bench.py --timesteps 64 --burnin 4 --steps 2 --warmup 1 --no-cpu --no-extras --profile-every 5 > gpurun_out/r11/b_dbg3.json 2> gpurun_out/r11/b_dbg3.err
echo done
