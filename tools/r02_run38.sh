cd $GRAFT_REPO_ROOT
for o in "pair=0" "no_tall=1" "max_splits=4" "max_splits=2" "no_tall=1 --mnv-opt max_splits=2" "no_tall=1 --mnv-opt max_splits=3"; do
echo "== $o"; timeout 300 python tools/opbench.py --filter matmult --mnv-opt $o --out gpurun_out/ob_a.json 2>&1 | grep "x256x\|x256 "
done
