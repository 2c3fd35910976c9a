set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r50
mkdir -p $O
P="--timesteps 8 --burnin 2 --steps 1 --warmup 1 --no-cpu --no-extras --profile-every 0"
python bench.py $P > $O/plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_mma_kernel" -s 4 -c 1 -o $O/prof python bench.py $P > $O/ncu.log 2>&1
echo done
