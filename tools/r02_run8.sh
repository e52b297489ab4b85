cd $GRAFT_REPO_ROOT
python tools/opbench.py --filter conv1 --out gpurun_out/ob_c1a.json 2>&1 | grep conv1
python tools/opbench.py --filter conv1 --mnv-opt s2d_im2col=1 --mnv-opt no_shift=1 --out gpurun_out/ob_c1b.json 2>&1 | grep conv1
python tools/opbench.py --filter conv1 --mnv-opt s2d_im2col=1 --out gpurun_out/ob_c1c.json 2>&1 | grep conv1
timeout 600 python bench.py --workload googlenet --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_goog.json 2> gpurun_out/r02_bench_goog.err; tail -c 300 gpurun_out/r02_bench_goog.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/r02_bench_goog.json'))
t=sum(v['ms_per_step'] for v in b['op_table'].values())
print("googlenet: sum op ms/step", t, "step", b['ms_per_step'], "launches/step", b['gpu_launches']/b['steps'], "calls/step", sum(v['calls_per_step'] for v in b['op_table'].values()))
for k,v in b['op_table'].items(): print("%-34s %6.1f %8.4f %s" % (k, v['calls_per_step'], v['ms_per_step'], round(v.get('frac',0),3)))
PY
