set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r19
mkdir -p $O
DCLL_WG2_PAIR=0 timeout 300 python tools/wgrad_err.py > $O/err_nopair.txt 2>&1
DCLL_WG2_PAIR=1 timeout 300 python tools/wgrad_err.py > $O/err_pair.txt 2>&1
echo done
