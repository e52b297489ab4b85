cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_b_gemm_conv.py tests/test_gpu_c_net.py -q -m gpu -x -k "twin" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_f_fullsize.py -q -m gpu -k "step" 2>&1 | grep -E "AssertionError|passed|failed|Error" | head
echo "== alexnet"; python bench.py --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'])"
echo "== googlenet"; python bench.py --workload googlenet --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'])"
