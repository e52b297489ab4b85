cd $GRAFT_REPO_ROOT
for w in googlenet alexnet lenet mlp; do python tools/host_profile.py $w 10 2>&1 | grep -E "enqueue"; done
