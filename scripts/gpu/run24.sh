cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r24_pytest.log 2>&1; tail -4 gpurun_out/r24_pytest.log | cut -c1-300
DIAG_REPS=30 timeout 300 python scripts/tc_diag.py 2>&1 | awk '{print $4, $5, $6, $7, $8, $9, $10}' | sort | uniq -c
