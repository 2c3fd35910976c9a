set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r53
mkdir -p $O
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 600 python -m pytest tests/test_dp_nccl.py -m gpu -q -x 2>&1 | tail -8) > $O/pytest_dp.log 2>&1
timeout 300 python bench.py --gpus 1 $Q > $O/b_n1.json 2> $O/b_n1.err
timeout 600 $TR --master-port 29511 bench.py --gpus 2 $Q > $O/b_n2.json 2> $O/b_n2.err
echo done
