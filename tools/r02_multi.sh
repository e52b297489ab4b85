cd $GRAFT_REPO_ROOT
for n in ${NG:-8}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_${n}gpu.json 2> gpurun_out/r02_bench_${n}gpu.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_${n}gpu.err
python - <<PY
import json
b=json.load(open('gpurun_out/r02_bench_${n}gpu.json'))
print("N=$n value", b['value'], "ms", b['ms_per_step'], "e2e", b['e2e']['value'], "f32", b['e2e']['f32_feed']['value'], "merge", b.get('merge_check'))
for k,v in b['other_configs'].items(): print(k, v.get('value'), v.get('ms_per_step'), (v.get('e2e') or {}).get('value'), v.get('merge_check',{}).get('ok'), v.get('error'))
PY
done
