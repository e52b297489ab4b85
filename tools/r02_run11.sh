cd $GRAFT_REPO_ROOT
for spec in "umma_gemm_kernel conv4_wgrad r02_conv4_wgrad" "umma_gemm_kernel conv2_bwd r02_conv2_dgrad" "umma_gemm_kernel conv3_fwd r02_conv3_fwd" "umma_gemm_kernel conv2_wgrad r02_conv2_wgrad" "maxpool332_bwd_idx pool1_bwd_idx r02_pool1_bwd_idx"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s 2 -c 1 -o gpurun_out/$3 -f python tools/one_op.py $2 4 > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail
