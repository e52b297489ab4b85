cd $GRAFT_REPO_ROOT
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'f32', d['e2e']['f32_feed']['value'], 'blocking', d['e2e']['blocking_read_value'], d['roofline']['achieved'], d['roofline']['frac'])"
done
