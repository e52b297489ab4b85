cd $GRAFT_REPO_ROOT
for i in 1 2; do
python bench.py --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['eager'], d['roofline']['achieved'], d['clocks'])"
done
