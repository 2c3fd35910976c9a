set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r55
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_f16x2.py tests/test_gpu_parity.py tests/test_gpu_readout.py -m gpu -x -q > $O/pytest.log 2>&1
tail -3 $O/pytest.log
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > $O/b.json 2> $O/b.err
P="--timesteps 8 --burnin 2 --steps 1 --warmup 1 --no-cpu --no-extras --profile-every 0"
python bench.py $P > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file $O/launches.csv python bench.py $P > $O/ncu_l.log 2>&1
echo done
