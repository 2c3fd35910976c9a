set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r22
mkdir -p $O
Q="--timesteps 256 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py --workload radio_ml_conv_train_16x16_B512_arp $Q > $O/b_16_512.json 2> $O/b_16_512.err
timeout 300 python bench.py --workload radio_ml_conv_train_16x16_B64 $Q > $O/b_16_64.json 2> $O/b_16_64.err
timeout 300 python bench.py --workload radio_ml_conv_train_16x16_B1024 $Q > $O/b_16_1024.json 2> $O/b_16_1024.err
Q="--timesteps 256 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 0"
timeout 300 python bench.py --workload radio_ml_conv_train_16x16_B512_arp $Q > $O/b_16_512_np.json 2> $O/b_16_512_np.err
timeout 300 python bench.py --workload radio_ml_conv_train_16x16_B64 $Q > $O/b_16_64_np.json 2> $O/b_16_64_np.err
echo done
