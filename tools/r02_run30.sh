cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_b_gemm_conv.py -q -m gpu -x 2>&1 | tail -5
echo "== pair"; timeout 300 python tools/opbench.py --filter matmult --out gpurun_out/ob_a.json 2>&1 | grep matmult
timeout 300 python tools/opbench.py --filter conv --out gpurun_out/ob_a.json 2>&1 | grep conv
echo "== pair=0"; timeout 300 python tools/opbench.py --filter matmult --mnv-opt pair=0 --out gpurun_out/ob_b.json 2>&1 | grep matmult
timeout 300 python tools/opbench.py --filter conv --mnv-opt pair=0 --out gpurun_out/ob_b.json 2>&1 | grep conv
