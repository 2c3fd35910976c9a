set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r17
mkdir -p $O
timeout 300 python tools/timeline.py --timesteps 64 > $O/tl_default.txt 2>&1
DCLL_TIMELINE_LAYER=1 timeout 300 python tools/timeline.py --timesteps 64 > $O/tl_default_l1.txt 2>&1
Q="--timesteps 128 --steps 3 --warmup 2 --no-cpu --no-extras --profile-every 7"
DCLL_PDL=0 timeout 300 python bench.py $Q > $O/b_nopdl.json 2> $O/b_nopdl.err
echo done
