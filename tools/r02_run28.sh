cd $GRAFT_REPO_ROOT
echo "== mn3"; python tools/opbench.py --filter matmult --out gpurun_out/ob_a.json 2>&1 | grep matmult
echo "== no_mn3"; python tools/opbench.py --filter matmult --mnv-opt no_mn3=1 --out gpurun_out/ob_b.json 2>&1 | grep matmult
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
