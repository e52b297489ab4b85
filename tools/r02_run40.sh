cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_c_net.py -q -m gpu -x 2>&1 | tail -5
timeout 300 python tools/e2e_probe.py
