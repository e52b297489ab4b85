set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 400 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference.json
timeout 600 python tools/opbench.py --out gpurun_out/opbench_final.json > gpurun_out/opbench_final.log 2>&1; tail -5 gpurun_out/opbench_final.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 500 -c 300 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
for spec in "conv_shift_fwd_kernel conv1_fwd shift_conv1_fwd" "umma_gemm_kernel conv3_fwd umma_conv3_fwd" "umma_gemm_kernel conv2_wgrad umma_conv2_wgrad" "lrn_bwd5_lite lrn1_bwd_lite lrn1_bwd_lite" "lrn_fwd5 lrn1_fwd_lite lrn1_fwd_lite" "maxpool332_bwd_idx pool1_bwd_idx pool1_bwd_idx" "s2d_nhwc conv1_fwd s2d_conv1"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s 2 -c 1 -o gpurun_out/final_$3 -f python tools/one_op.py $2 4 > /dev/null 2>&1
done
ls -la gpurun_out | grep final_
