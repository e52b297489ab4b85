cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_b_gemm_conv.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python tools/opbench.py --filter conv --out gpurun_out/ob_a.json 2>&1 | grep "forward\|backward_data"
for w in alexnet googlenet; do
echo "== $w"; timeout 600 python bench.py --workload $w --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['eager']['ms_per_step'])"
done
