set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
(time python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60) > gpurun_out/r1/pytest.log 2>&1
python __graft_entry__.py --smoke > gpurun_out/r1/smoke.log 2>&1; echo smoke rc=$? >> gpurun_out/r1/smoke.log
W=radio_ml_conv_train_128x128_B64
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
python bench.py $Q > gpurun_out/r1/b_default.json 2> gpurun_out/r1/b_default.err
DCLL_CONV_MMA2_STAGES=4 python bench.py $Q > gpurun_out/r1/b_stages4.json 2> gpurun_out/r1/b_stages4.err
DCLL_TRACE_FUSE=0 python bench.py $Q > gpurun_out/r1/b_nofuse.json 2> gpurun_out/r1/b_nofuse.err
DCLL_TRACE_FUSE=0 DCLL_CONV_MMA2=3 python bench.py $Q > gpurun_out/r1/b_nofuse_mma2all.json 2> gpurun_out/r1/b_nofuse_mma2all.err
DCLL_TRACE_FUSE=0 DCLL_CONV_MMA2=3 DCLL_CONV_MMA2_STAGES=4 python bench.py $Q > gpurun_out/r1/b_nofuse_mma2all_s4.json 2> gpurun_out/r1/b_nofuse_mma2all_s4.err
DCLL_CONV_MMA2=3 python bench.py $Q > gpurun_out/r1/b_fuse_mma2all.json 2> gpurun_out/r1/b_fuse_mma2all.err
(time python bench.py --steps 3 --warmup 3) > gpurun_out/r1/b_full.json 2> gpurun_out/r1/b_full.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1/b_ref.json 2> gpurun_out/r1/b_ref.err
echo done
