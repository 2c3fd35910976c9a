set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r33
mkdir -p $O
(timeout 900 python -m pytest tests/test_dp_nccl.py -m gpu -q -x 2>&1 | tail -8) > $O/pytest_dp.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras > $O/b_n2_f16.json 2> $O/b_n2_f16.err
timeout 600 python bench.py --timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras > $O/b_n1_f16.json 2> $O/b_n1_f16.err
echo done
