cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_b_gemm_conv.py -q -m gpu -x 2>&1 | tail -3
for w in alexnet googlenet; do for f in "" "--no-pdl"; do
echo "== $w $f"; timeout 600 python bench.py --workload $w $f --no-e2e --no-cpu-baseline --no-other-configs 2>gpurun_out/err_$w.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['eager']['ms_per_step'])"; tail -3 gpurun_out/err_$w.txt
done; done
