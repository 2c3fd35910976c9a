set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r6
nvidia-smi -L
(time timeout 900 python -m pytest tests/test_dp_nccl.py -m gpu -q -x -s 2>&1 | tail -30) > gpurun_out/r6/pytest_dp.log 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python bench.py --gpus 1 $Q > gpurun_out/r6/b_n1.json 2> gpurun_out/r6/b_n1.err
timeout 600 $TR --master-port 29511 bench.py --gpus 2 $Q > gpurun_out/r6/b_n2.json 2> gpurun_out/r6/b_n2.err
DCLL_DP_MAX_CTAS=4 timeout 600 $TR --master-port 29512 bench.py --gpus 2 $Q > gpurun_out/r6/b_n2_cta4.json 2> gpurun_out/r6/b_n2_cta4.err
DCLL_DP_MAX_CTAS=0 timeout 600 $TR --master-port 29513 bench.py --gpus 2 $Q > gpurun_out/r6/b_n2_cta0.json 2> gpurun_out/r6/b_n2_cta0.err
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --workload radio_ml_conv_train_dp_global8192_16x16 --timesteps 256 --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r6/b_g8192_16_n2.json 2> gpurun_out/r6/b_g8192_16_n2.err
echo done
