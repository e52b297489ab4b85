cd $GRAFT_REPO_ROOT
echo "== mn3"; python tools/opbench.py --filter conv --out gpurun_out/ob_a.json 2>&1 | grep backward_filter
echo "== no_mn3"; python tools/opbench.py --filter conv --mnv-opt no_mn3=1 --out gpurun_out/ob_b.json 2>&1 | grep backward_filter
