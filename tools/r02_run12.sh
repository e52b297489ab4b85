cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_b_gemm_conv.py -q -m gpu -x 2>&1 | tail -5
python tools/opbench.py --filter conv --out gpurun_out/ob_conv_ml.json 2>&1 | grep conv
python tools/opbench.py --filter matmult --out gpurun_out/ob_mm_ml.json 2>&1 | grep matmult
