cd $GRAFT_REPO_ROOT
for o in "tall_min_stages=32" "no_tall=1" "tall_min_stages=8" "tall_min_stages=64" "no_tail=1" "no_transposed=1"; do
echo "== $o"; python bench.py --mnv-opt $o --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['eager']['ms_per_step'], d['roofline']['achieved'])"
done
