cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DIAG_REPS=30 timeout 150 python scripts/tc_diag.py 2>&1 | grep -v "kernel added\|bad rows" | cut -c1-160 > gpurun_out/r44_diag_shipped.txt; grep -c "bad frac 0.0 " gpurun_out/r44_diag_shipped.txt; grep -v "bad frac 0.0 " gpurun_out/r44_diag_shipped.txt | head -6
