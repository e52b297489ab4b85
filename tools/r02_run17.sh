cd $GRAFT_REPO_ROOT
for opt in "" "--mnv-opt force_tma_a=1"; do
timeout 600 python bench.py --workload googlenet --no-cpu-baseline --no-e2e $opt > gpurun_out/r02_goog_x.json 2> gpurun_out/r02_goog_x.err; tail -c 200 gpurun_out/r02_goog_x.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r02_goog_x.json'))
print(b['config'].get('tuning'), b['value'], b['ms_per_step'])
for k,v in list(b['op_table'].items())[:6]: print("  ", k, {a:(round(x,4) if isinstance(x,float) else x) for a,x in v.items() if a in ('calls_per_step','ms_per_step','frac')})
PY
done
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-other-configs --mnv-opt force_tma_a=1 > gpurun_out/r02_alex_x.json 2>/dev/null; python -c "
import json; b=json.load(open('gpurun_out/r02_alex_x.json')); print('alexnet force_tma_a', b['value'], b['ms_per_step'])"
