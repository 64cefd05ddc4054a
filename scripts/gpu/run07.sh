cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q -x -k "test_tc_clip_loss" 2>&1 | tail -25 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_fullmodel.py -q -x -k "clip_loss_large" 2>&1 | tail -15 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "clip_loss" 2>&1 | tail -5 | cut -c1-300
