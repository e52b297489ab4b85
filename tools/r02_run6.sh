cd $GRAFT_REPO_ROOT
python tools/opbench.py --filter conv --out gpurun_out/ob_new.json 2>&1 | grep filter
python tools/opbench.py --filter conv --mnv-opt no_nhwc_wgrad=1 --out gpurun_out/ob_old.json 2>&1 | grep filter
