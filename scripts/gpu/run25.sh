# compute-sanitizer on the kernels added / changed in this session: two-group tcgen05 GEMM epilogue, cp.async attention staging,
# fused ConvMixer stages (last-CTA reductions), closed-form attention pooling, masked MSE
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SEL="test_tc_linear_fwd or test_tc_attention_fwd_bwd or test_convmixer_shapes_vs_oracle or test_attn_pool_closed_form or test_masked_lc or test_convmixer_dropout"
for tool in racecheck synccheck memcheck; do
  timeout 1500 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_round2.py -q -k "$SEL" > gpurun_out/r25_sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "passed|failed|SUMMARY" gpurun_out/r25_sanitizer_$tool.log | tail -3
done
