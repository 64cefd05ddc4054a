cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python scripts/tc_diag.py > gpurun_out/r38_diag.txt 2>&1; cat gpurun_out/r38_diag.txt | cut -c1-250
