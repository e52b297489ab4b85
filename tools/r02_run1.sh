set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 1500 python -m pytest tests -q -m gpu --maxfail=15 -s > gpurun_out/r02_pytest1.log 2>&1; tail -40 gpurun_out/r02_pytest1.log
timeout 300 python bench.py > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; tail -c 600 gpurun_out/r02_bench1.err; head -c 600 gpurun_out/r02_bench1.json
timeout 400 python tools/opbench.py --out gpurun_out/r02_opbench1.json > gpurun_out/r02_opbench1.log 2>&1; tail -5 gpurun_out/r02_opbench1.log
