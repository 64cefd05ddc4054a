cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q 2>&1 | tail -15 | cut -c1-300
# sanitizer evidence for the tcgen05 kernels (VERDICT weak #4): racecheck + synccheck on the aux-ring GEMM cases and the weight-gradient kernel
for tool in racecheck synccheck memcheck; do
timeout 420 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_tc.py -q -x -k "test_tc_linear_bwd_input or test_tc_linear_res_ln or test_tc_linear_bwd_weight" > gpurun_out/r06_sanitizer_$tool.log 2>&1
echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r06_sanitizer_$tool.log | tail -5 | cut -c1-250
done
timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_fused.py -q -x -k "test_ffn_fused_fwd_bwd and 4100" > gpurun_out/r06_sanitizer_racecheck_fused.log 2>&1
echo "== racecheck fused rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r06_sanitizer_racecheck_fused.log | tail -5 | cut -c1-250
