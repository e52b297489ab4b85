cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_b_gemm_conv.py -q -m gpu -x 2>&1 | tail -5
for o in "pair=1" "pair=0"; do
echo "== $o"; timeout 300 python tools/opbench.py --filter matmult --mnv-opt $o --out gpurun_out/ob_a.json 2>&1 | grep "8192\|9216x256"
timeout 300 python tools/opbench.py --filter conv --mnv-opt $o --out gpurun_out/ob_a.json 2>&1 | grep "conv4\|conv2"
done
