cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu --maxfail=15 -s > gpurun_out/r02_pytest5.log 2>&1; tail -25 gpurun_out/r02_pytest5.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err; tail -c 800 gpurun_out/r02_bench5.err; python - <<'PY'
import json
try:
    b=json.load(open('gpurun_out/r02_bench5.json'))
    print(b['value'], b['ms_per_step'], b['e2e']['value'], b['roofline']['achieved'], b['roofline']['frac'])
    for k,v in b['other_configs'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
    for k,v in list(b['op_table'].items())[:12]: print(k, {a:(round(x,4) if isinstance(x,float) else x) for a,x in v.items()})
except Exception as e: print("ERR",e)
PY
