set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r47
mkdir -p $O
(time python bench.py --steps 5 --warmup 3) > $O/b_full.json 2> $O/b_full.err
echo done
