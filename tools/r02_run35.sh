cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_a_memops.py tests/test_gpu_c_net.py tests/test_gpu_f_fullsize.py -q -m gpu -x -k "add_n or step or graph or fusion or googlenet" 2>&1 | tail -5
for w in googlenet alexnet; do
echo "== $w"; timeout 600 python bench.py --workload $w --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['eager']['ms_per_step'])"
done
