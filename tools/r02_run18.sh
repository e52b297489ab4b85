cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_b_gemm_conv.py tests/test_gpu_f_fullsize.py -q -m gpu -x 2>&1 | tail -6
echo "== transposed"; python tools/opbench.py --filter conv2 --mnv-opt no_tail=0 --out gpurun_out/ob_c2_t.json 2>&1 | grep conv2
echo "== normal"; python tools/opbench.py --filter conv2 --mnv-opt no_transposed=1 --out gpurun_out/ob_c2_n.json 2>&1 | grep conv2
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r02_bench18.json 2> gpurun_out/r02_bench18.err; tail -c 300 gpurun_out/r02_bench18.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r02_bench18.json'))
print(b['value'], b['ms_per_step'], b['roofline']['achieved'], b['roofline']['frac'])
for k,v in b['other_configs'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
for k,v in list(b['op_table'].items())[:4]: print(k, {a:(round(x,4) if isinstance(x,float) else x) for a,x in v.items()})
PY
