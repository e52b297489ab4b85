cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_b_gemm_conv.py tests/test_gpu_c_net.py -q -m gpu -x 2>&1 | tail -3
for w in alexnet googlenet; do
echo "== $w"; timeout 600 python bench.py --workload $w --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['eager']['ms_per_step'])
t=d['op_table']
for k in ('mnv_relu_backward_tw','mnv_conv_forward_tw','mnv_conv_backward_filter_tw','mnv_conv_backward_data_tw'):
    print(' ', k, round(t[k]['ms_per_step'],3), t[k].get('gbs') or t[k].get('tflops'))"
done
