set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r43
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_f16x2.py tests/test_gpu_headline.py tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1
tail -5 $O/pytest.log
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > $O/b.json 2> $O/b.err
P="--timesteps 8 --burnin 2 --steps 1 --warmup 1 --no-cpu --no-extras --profile-every 0"
python bench.py $P > $O/plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel|conv_mma_kernel|trace_image1_kernel" -s 12 -c 6 -o $O/prof python bench.py $P > $O/ncu.log 2>&1
echo done
