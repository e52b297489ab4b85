cd $GRAFT_REPO_ROOT
python tools/opbench.py --filter conv --out gpurun_out/ob_new.json 2>&1 | grep filter
python tools/opbench.py --filter conv --mnv-opt no_nhwc_wgrad=1 --out gpurun_out/ob_old.json 2>&1 | grep filter
timeout 600 python -m pytest tests/test_gpu_a_memops.py -q -m gpu -k sgd 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-other-configs > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err; tail -c 500 gpurun_out/r02_bench6.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r02_bench6.json'))
print(b['value'], b['ms_per_step'], b['e2e']['value'], b['roofline']['achieved'], b['roofline']['frac'])
for k,v in list(b['op_table'].items())[:8]: print(k, {a:(round(x,4) if isinstance(x,float) else x) for a,x in v.items()})
PY
