cd $GRAFT_REPO_ROOT
for m in enqueue event blocking; do for n in lenet mlp; do ./minerva_b200/lib/mnist_apps --net $n --completion $m --steps 200 --warmup 20; done; done
timeout 900 python -m pytest tests/test_gpu_d_cpp_plugin.py tests/test_gpu_a_memops.py -q -m gpu -k "cpp or sgd or exact" 2>&1 | tail -15
python tools/opbench.py --filter conv --out gpurun_out/ob_new.json 2>&1 | grep filter
python tools/opbench.py --filter conv --mnv-opt no_nhwc_wgrad=1 --out gpurun_out/ob_old.json 2>&1 | grep filter
timeout 600 python bench.py --no-cpu-baseline --no-other-configs > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; tail -c 500 gpurun_out/r02_bench7.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r02_bench7.json'))
print(b['value'], b['ms_per_step'], b['e2e']['value'], b['roofline']['achieved'], b['roofline']['frac'])
for k,v in list(b['op_table'].items())[:8]: print(k, {a:(round(x,4) if isinstance(x,float) else x) for a,x in v.items()})
PY
