cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_f_fullsize.py tests/test_gpu_b_gemm_conv.py -q -m gpu --maxfail=15 -s -k "step_matches or tail_split" > gpurun_out/r02_pytest2.log 2>&1; tail -30 gpurun_out/r02_pytest2.log | cut -c1-400
