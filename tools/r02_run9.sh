cd $GRAFT_REPO_ROOT
for t in 64 32 16 8; do echo "== tall_min_stages=$t"; python tools/opbench.py --filter matmult --mnv-opt tall_min_stages=$t --out gpurun_out/ob_mm_t$t.json 2>&1 | grep matmult; done
echo "== conv tall 32"; python tools/opbench.py --filter conv --mnv-opt tall_min_stages=32 --out gpurun_out/ob_conv_t32.json 2>&1 | grep conv
echo "== conv tall 16"; python tools/opbench.py --filter conv --mnv-opt tall_min_stages=16 --out gpurun_out/ob_conv_t16.json 2>&1 | grep conv
