set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r30
mkdir -p $O
P="--timesteps 6 --burnin 2 --steps 1 --warmup 1 --no-cpu --no-extras --profile-every 0 --precision f16x2"
python bench.py $P > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc2p_kernel|conv_mma2_kernel" -s 4 -c 3 -o $O/prof_f16 python bench.py $P > $O/ncu.log 2>&1
echo done
