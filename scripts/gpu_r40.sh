cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 100 python scripts/tc_diag.py 2>&1 | grep -v "kernel added\|bad rows" | cut -c1-200 | tee gpurun_out/r40_diag.txt
