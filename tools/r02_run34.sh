cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_b_gemm_conv.py -q -m gpu -x 2>&1 | tail -8
timeout 300 python tools/im2col_probe.py
for f in "" "--mnv-opt no_pointwise=1"; do
echo "== googlenet $f"; timeout 600 python bench.py --workload googlenet $f --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['eager']['ms_per_step'])"
done
