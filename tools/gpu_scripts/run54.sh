set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r54
mkdir -p $O
(time python bench.py --steps 5 --warmup 3) > $O/b_full.json 2> $O/b_full.err
P="--timesteps 8 --burnin 2 --steps 1 --warmup 1 --no-cpu --no-extras --profile-every 0"
python bench.py $P > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file $O/launches.csv python bench.py $P > $O/ncu_l.log 2>&1
python bench.py $P > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc2p_kernel|wgrad_tc_kernel|conv_mma2_kernel|conv_mma_kernel|trace_image|readout_bwd2_kernel|readout_tc_kernel|wout_grad_adam2|readout_finish" -s 70 -c 24 -o $O/prof python bench.py $P > $O/ncu.log 2>&1
echo done
