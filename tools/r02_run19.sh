cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_b_gemm_conv.py -q -m gpu -x -k "operand_paths or relu_fused or three" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_f_fullsize.py -q -m gpu -k "googlenet_step" 2>&1 | grep -E "AssertionError|passed|failed" | head
echo "== transposed v4"; python tools/opbench.py --filter conv2 --out gpurun_out/ob_c2_t.json 2>&1 | grep conv2
