cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_tc.py -q -k "attention or seq_encoder" 2>&1 | tail -4 | cut -c1-300
python scripts/bench_fused.py attn 2>&1 | tail -2
