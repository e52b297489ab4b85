cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_f_fullsize.py tests/test_gpu_a_memops.py -q -m gpu -x -k "concat or googlenet" 2>&1 | tail -3
echo "== googlenet"; timeout 600 python bench.py --workload googlenet --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['eager']['ms_per_step'])"
