set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r38
mkdir -p $O
Q="--timesteps 64 --burnin 4 --steps 2 --warmup 1 --no-cpu --no-extras --profile-every 5"
for d in 0 1 2 4 3 7; do
DCLL_RB2_DEBUG=$d timeout 300 python bench.py $Q > $O/b_dbg$d.json 2> $O/b_dbg$d.err
done
echo done
