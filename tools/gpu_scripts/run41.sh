set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r41
mkdir -p $O
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
DCLL_RB2_SLICES=1 timeout 300 python bench.py $Q > $O/b_s1.json 2> $O/b_s1.err
DCLL_RB2_SLICES=4 timeout 300 python bench.py $Q > $O/b_s4.json 2> $O/b_s4.err
timeout 300 python bench.py $Q > $O/b_s2.json 2> $O/b_s2.err
echo done
