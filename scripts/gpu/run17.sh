cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_tc.py -q -k "clip_loss" 2>&1 | tail -4 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_fullmodel.py -q -k "clip_loss_large" 2>&1 | tail -4 | cut -c1-300
python scripts/bench_fused.py loss 2>&1 | tail -6
