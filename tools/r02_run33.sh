cd $GRAFT_REPO_ROOT
for d in . _prev . _prev; do
(cd $d && python bench.py --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$d', d['value'], d['ms_per_step'], d['eager']['ms_per_step'], d['roofline']['achieved'])")
done
