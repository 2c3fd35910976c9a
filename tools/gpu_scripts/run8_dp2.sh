set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r8
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python bench.py --gpus 1 $Q > gpurun_out/r8/b_n1.json 2> gpurun_out/r8/b_n1.err
timeout 600 $TR --master-port 29511 bench.py --gpus 2 $Q > gpurun_out/r8/b_n2_cta8.json 2> gpurun_out/r8/b_n2_cta8.err
DCLL_DP_MAX_CTAS=4 timeout 600 $TR --master-port 29512 bench.py --gpus 2 $Q > gpurun_out/r8/b_n2_cta4.json 2> gpurun_out/r8/b_n2_cta4.err
DCLL_DP_MAX_CTAS=2 timeout 600 $TR --master-port 29513 bench.py --gpus 2 $Q > gpurun_out/r8/b_n2_cta2.json 2> gpurun_out/r8/b_n2_cta2.err
(timeout 600 python -m pytest tests/test_dp_nccl.py tests/test_gpu_parity.py -m gpu -q -x -k "dp or image2spike" 2>&1 | tail -8) > gpurun_out/r8/pytest.log 2>&1
echo done
