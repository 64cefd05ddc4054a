cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 160 -c 8 -o gpurun_out/r34_gemm python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r34_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad -s 80 -c 4 -o gpurun_out/r34_wgrad python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r34_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
