cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo rc=$?; tail -c 1500 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
try:
    b=json.load(open('gpurun_out/r02_bench_2gpu.json'))
    for k in b:
        if k not in ('op_table','cpu_reference_ops'): print(k, json.dumps(b[k])[:900])
except Exception as e: print("ERR",e)
PY
timeout 600 python -m pytest tests/test_gpu_e_multi.py tests/test_gpu_c_net.py -q -m gpu -s 2>&1 | tail -15
