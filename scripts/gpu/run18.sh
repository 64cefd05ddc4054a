cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -k "attention or seq_encoder" 2>&1 | tail -3 | cut -c1-300
for i in 1 2; do python scripts/bench_fused.py attn 2>&1 | tail -2; done
