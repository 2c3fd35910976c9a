set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r10
Q="--timesteps 64 --burnin 4 --steps 2 --warmup 1 --no-cpu --no-extras --profile-every 5"
for d in 0 1 2 3 4 5 7; do
DCLL_CONV_DEBUG=$d timeout 300 python bench.py $Q > gpurun_out/r10/b_dbg$d.json 2> gpurun_out/r10/b_dbg$d.err
done
for d in 0 1 4; do
DCLL_CONV_MMA2=3 DCLL_TRACE_FUSE=0 DCLL_CONV_DEBUG=$d timeout 300 python bench.py $Q > gpurun_out/r10/b_all2_dbg$d.json 2> gpurun_out/r10/b_all2_dbg$d.err
done
echo done
