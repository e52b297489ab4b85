# round-2 final artefacts (1 GPU): GPU test suite, bench lines, op bench, ncu launch lists, full captures of the dominant kernels
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $O/r02_pytest_gpu.log 2>&1; tail -3 $O/r02_pytest_gpu.log
timeout 900 python bench.py > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; head -c 400 $O/r02_bench_1gpu.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; head -c 300 $O/r02_bench_reference.json; echo
timeout 600 python bench.py --impl reference --workload mlp --steps 3 --warmup 1 > $O/r02_bench_reference_mlp.json 2> $O/r02_bench_reference_mlp.err; head -c 300 $O/r02_bench_reference_mlp.json; echo
timeout 600 python bench.py --workload googlenet > $O/r02_bench_googlenet_1gpu.json 2> $O/r02_bench_googlenet_1gpu.err; head -c 300 $O/r02_bench_googlenet_1gpu.json; echo
timeout 600 python tools/opbench.py --out $O/r02_opbench.json > $O/r02_opbench.log 2>&1; tail -3 $O/r02_opbench.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none -s 400 -c 260 --csv --log-file $O/r02_launches_bench.csv python bench.py --graph off --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs > $O/r02_bench_under_ncu.json 2> $O/r02_bench_under_ncu.err
python tools/launch_summary.py $O/r02_launches_bench.csv $O/r02_launch_list_summary.json > $O/r02_launch_list_summary.txt; head -12 $O/r02_launch_list_summary.txt
timeout 1200 ncu --metrics $M --clock-control none -s 2900 -c 800 --csv --log-file $O/r02_launches_googlenet.csv python bench.py --workload googlenet --graph off --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs > $O/r02_goog_under_ncu.json 2> $O/r02_goog_under_ncu.err
python tools/launch_summary.py $O/r02_launches_googlenet.csv $O/r02_launch_list_googlenet.json > $O/r02_launch_list_googlenet.txt; head -8 $O/r02_launch_list_googlenet.txt
for spec in "umma_gemm_kernel conv4_wgrad umma_conv4_wgrad" "umma_gemm_kernel conv2_wgrad umma_conv2_wgrad" "umma_gemm_kernel conv3_fwd umma_conv3_fwd" "umma_gemm_kernel conv4_bwd umma_conv4_dgrad" "nchw_to_nhwc conv3_fwd nhwc_conv3"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s 2 -c 1 -o $O/r02_full_$3 -f python tools/one_op.py $2 4 > /dev/null 2>&1
done
ls -la $O | grep r02_full_
