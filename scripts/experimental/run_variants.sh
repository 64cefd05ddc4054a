# on the GPU box: failures per variant (each line of tc_diag output without "bad frac 0.0" is a failing run)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in scripts/experimental/lib_*.so; do
  v=$(basename $lib .so)
  MVN_DIAG_LIB=$GRAFT_REPO_ROOT/$lib DIAG_REPS=${DIAG_REPS:-10} timeout 150 python scripts/tc_diag.py > gpurun_out/variants_$v.log 2>&1
  grep "bad frac" gpurun_out/variants_$v.log > gpurun_out/variants_$v.txt
  echo "$v: $(grep -vc 'bad frac 0.0 ' gpurun_out/variants_$v.txt) failing of $(wc -l < gpurun_out/variants_$v.txt)"
  [ -s gpurun_out/variants_$v.txt ] || tail -5 gpurun_out/variants_$v.log | cut -c1-300
done
