cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_a_memops.py tests/test_gpu_f_fullsize.py -q -m gpu -k "pool" 2>&1 | tail -4
echo "== flat"; python tools/opbench.py --filter pool --out gpurun_out/ob_pool_flat.json 2>&1 | grep "bwd_idx"
echo "== block"; MNV_POOL_STRIP=0 python tools/opbench.py --filter pool --out gpurun_out/ob_pool_block.json 2>&1 | grep "bwd_idx"
