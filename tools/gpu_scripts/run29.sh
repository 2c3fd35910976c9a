set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r29
mkdir -p $O
timeout 120 tools/_build/mma_bench > $O/mma_bench.txt 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py --precision f16x2 $Q > $O/b_f16.json 2> $O/b_f16.err
timeout 300 python bench.py $Q > $O/b_bf16.json 2> $O/b_bf16.err
echo done
