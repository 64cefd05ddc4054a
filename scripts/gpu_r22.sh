set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_mma -s 20 -c 2 -o gpurun_out/r22_attn_bwd python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r22_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_mma -s 20 -c 2 -o gpurun_out/r22_attn_fwd python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r22_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 200 -c 12 -o gpurun_out/r22_gemm python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r22_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad -s 100 -c 8 -o gpurun_out/r22_wgrad python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r22_ncu4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_bwd_vec -s 30 -c 4 -o gpurun_out/r22_lnbwd python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r22_ncu5.log 2>&1
ls -la gpurun_out/; du -sh gpurun_out
