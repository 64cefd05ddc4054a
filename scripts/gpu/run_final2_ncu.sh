# Final measurement pass, part 2: `ncu --set full` captures of every kernel class (two launches each).  The reports are exported to
# raw CSV on the box (gpurun brings back at most 64 MiB) and summarised on the CPU box by scripts/ncu_summary.py into profiles/.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=r02b_final
# --set full captures, two launches per kernel class
for k in ffn_bwd_kernel ffn_fwd_kernel attn_bwd_mma_kernel attn_fwd_mma_kernel tc_gemm_kernel tc_wgrad_kernel ln_bwd_vec_kernel embed_fwd_kernel tc_lse_kernel tc_grad_kernel; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -f -o gpurun_out/${T}_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweep --no-graph > /dev/null 2>&1
done
for k in patch_conv_fwd_kernel patch_conv_wgrad_kernel mixer_fwd_kernel mixer_bwd_kernel; do
timeout 300 ncu --set full --clock-control none -k regex:$k -s 2 -c 4 -f -o gpurun_out/${T}_$k python bench.py --steps 1 --warmup 3 --workload c3 --no-cpu-baseline --no-sweep --no-graph > /dev/null 2>&1
done
for r in gpurun_out/${T}_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; done
# keep the source-level pages of the two attention kernels (stall sampling per SASS instruction), drop the reports
for k in attn_fwd_mma_kernel attn_bwd_mma_kernel; do ncu -i gpurun_out/${T}_$k.ncu-rep --page source --csv --print-source sass > gpurun_out/${T}_$k.source.csv 2>/dev/null; done
rm -f gpurun_out/${T}_*.ncu-rep
ls -la gpurun_out | head -40; du -sh gpurun_out
