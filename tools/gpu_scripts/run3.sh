set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r3
Q="--timesteps 64 --burnin 4 --steps 2 --warmup 1 --no-cpu --no-extras --profile-every 5"
for d in 0 1 2 3 4 7; do
DCLL_WG2_DEBUG=$d timeout 300 python bench.py $Q > gpurun_out/r3/b_dbg$d.json 2> gpurun_out/r3/b_dbg$d.err
done
P="--timesteps 6 --burnin 2 --steps 1 --warmup 1 --no-cpu --no-extras --profile-every 0"
python bench.py $P > gpurun_out/r3/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc2_kernel|conv_mma2_kernel|conv_mma_kernel|trace_image_kernel|readout_bwd2_kernel|readout_tc_kernel|wout_grad_adam2" -s 60 -c 16 -o gpurun_out/r3/prof python bench.py $P > gpurun_out/r3/ncu.log 2>&1
echo done
