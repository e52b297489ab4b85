# round-2 profile artefacts: ncu launch list of the bench command (2 timed steps), summary, op bench, 1-GPU bench line
cd $GRAFT_REPO_ROOT
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 400 -c 260 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_bench_under_ncu.err
python tools/launch_summary.py gpurun_out/r02_launches_bench.csv gpurun_out/r02_launch_list_summary.json > gpurun_out/r02_launch_list_summary.txt; head -40 gpurun_out/r02_launch_list_summary.txt
timeout 600 python tools/opbench.py --out gpurun_out/r02_opbench.json > gpurun_out/r02_opbench.log 2>&1; tail -3 gpurun_out/r02_opbench.log
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; head -c 300 gpurun_out/r02_bench_1gpu.json
