cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_c_net.py tests/test_gpu_f_fullsize.py -q -m gpu --maxfail=15 -s -k "step_matches or image_transform or feed_data" > gpurun_out/r02_pytest3.log 2>&1; tail -30 gpurun_out/r02_pytest3.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; tail -c 1500 gpurun_out/r02_bench3.err; python - <<'PY'
import json
try:
    b=json.load(open('gpurun_out/r02_bench3.json'))
    for k in b:
        if k not in ('op_table',): print(k, json.dumps(b[k])[:700])
    for k,v in b['op_table'].items(): print(k, {a:(round(x,4) if isinstance(x,float) else x) for a,x in v.items()})
except Exception as e: print("ERR",e)
PY
