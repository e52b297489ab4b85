cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_a_memops.py tests/test_gpu_f_fullsize.py -q -m gpu -x -k "pool" 2>&1 | tail -3
echo "== googlenet"; timeout 600 python bench.py --workload googlenet --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['eager']['ms_per_step'])
t=d['op_table']
for k,v in sorted(t.items(), key=lambda kv:-kv[1]['ms_per_step'])[:14]:
    print(' ', k, v['calls_per_step'], round(v['ms_per_step'],3), v.get('gbs') and round(v['gbs']) or v.get('tflops') and round(v['tflops']))"
