cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in va vb; do echo "== $v"; MVN_DIAG_LIB=$GRAFT_REPO_ROOT/scripts/experimental/lib_$v.so timeout 100 python scripts/tc_diag.py 2>&1 | grep -v "kernel added\|bad rows" | cut -c1-160 > gpurun_out/r42_$v.txt; grep -c "bad frac 0.0 " gpurun_out/r42_$v.txt; grep -v "bad frac 0.0 " gpurun_out/r42_$v.txt | head -6; done
