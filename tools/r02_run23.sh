cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_c_net.py tests/test_gpu_a_memops.py -q -m gpu -x -k "graph or generators" 2>&1 | tail -15
for w in mlp lenet alexnet googlenet; do
echo "== $w"; timeout 600 python bench.py --workload $w --no-cpu-baseline --no-other-configs 2>gpurun_out/err_$w.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['host_enqueue_ms_per_step'], d['cuda_graph'], d['eager'])
print({k:v for k,v in d['e2e'].items() if k in ('value','f32_feed','blocking_read_value')})"; tail -3 gpurun_out/err_$w.txt
done
