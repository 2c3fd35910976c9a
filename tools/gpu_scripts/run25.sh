set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r25
mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_readout.py tests/test_gpu_tensorcore.py tests/test_gpu_headline.py -m gpu -q -x -k "not accuracy" 2>&1 | tail -8) > $O/pytest.log 2>&1
DCLL_PRECISION=f16x2 timeout 300 python tools/tc_check.py --wgrad > $O/check_f16.txt 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py --precision f16x2 $Q > $O/b_f16.json 2> $O/b_f16.err
timeout 300 python bench.py $Q > $O/b_bf16.json 2> $O/b_bf16.err
DCLL_PRECISION=f16x2 timeout 300 python tools/timeline.py --timesteps 32 > $O/tl_f16.txt 2>&1
echo done
