cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_b_gemm_conv.py -q -m gpu -x 2>&1 | tail -3
echo "== mn3"; python tools/opbench.py --filter add --out gpurun_out/ob_a.json 2>&1 | grep matmult
echo "== no_mn3"; python tools/opbench.py --filter add --mnv-opt no_mn3=1 --out gpurun_out/ob_b.json 2>&1 | grep matmult
python bench.py --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['eager']['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])"
