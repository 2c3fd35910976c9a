set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r36
mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -q -x -k "not accuracy" 2>&1 | tail -8) > $O/pytest.log 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
timeout 300 python bench.py $Q > $O/b_f16.json 2> $O/b_f16.err
DCLL_SPIKE_PACK=0 timeout 300 python bench.py $Q > $O/b_f16_nopack.json 2> $O/b_f16_nopack.err
timeout 300 python bench.py --precision bf16x3 $Q > $O/b_bf16.json 2> $O/b_bf16.err
echo done
