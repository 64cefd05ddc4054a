cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for k in attn_fwd_mma_kernel attn_bwd_mma_kernel; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 2 -f -o gpurun_out/r23_$k python scripts/bench_fused.py attn > /dev/null 2>&1
done
ls -la gpurun_out/r23_*
