cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_a_memops.py -q -m gpu -k "pooling_bit_exact and case9" 2>&1 | grep -E "Error|assert|differ|fwd|bwd" | head -20
