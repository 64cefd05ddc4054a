cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/bench_fused.py attn 2>&1 | tail -2
for k in ffn_bwd_kernel ffn_fwd_kernel; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -f -o gpurun_out/r03_$k python scripts/bench_fused.py ffn > gpurun_out/r03_ncu_$k.log 2>&1; tail -2 gpurun_out/r03_ncu_$k.log | cut -c1-200
done
for k in attn_bwd_mma_kernel attn_fwd_mma_kernel; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -f -o gpurun_out/r03_$k python scripts/bench_fused.py attn > gpurun_out/r03_ncu_$k.log 2>&1; tail -2 gpurun_out/r03_ncu_$k.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
