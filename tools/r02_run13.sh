cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_a_memops.py tests/test_gpu_f_fullsize.py tests/test_gpu_c_net.py -q -m gpu -k "pool or net or step or fusion" 2>&1 | tail -8
echo "== strip"; python tools/opbench.py --filter pool --out gpurun_out/ob_pool_strip.json 2>&1 | grep pool
echo "== block"; MNV_POOL_STRIP=0 python tools/opbench.py --filter pool --out gpurun_out/ob_pool_block.json 2>&1 | grep pool
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r02_bench13.json 2> gpurun_out/r02_bench13.err; tail -c 300 gpurun_out/r02_bench13.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r02_bench13.json'))
print(b['value'], b['ms_per_step'], b['roofline']['achieved'], b['roofline']['frac'])
for k,v in b['other_configs'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
for k,v in list(b['op_table'].items())[:12]: print(k, {a:(round(x,4) if isinstance(x,float) else x) for a,x in v.items()})
PY
