cd $GRAFT_REPO_ROOT
for w in alexnet googlenet lenet mlp; do
echo "== $w"; python bench.py --workload $w --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['host_enqueue_ms_per_step'])"
done
