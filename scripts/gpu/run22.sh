cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "attn_pool or enc_lc_attn or masked_lc or seq_encoder_golden" 2>&1 | grep -E "Error|error|passed|failed|assert" | tail -12 | cut -c1-300
