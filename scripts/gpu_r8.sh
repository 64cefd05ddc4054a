set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -s -k "attention or seq_encoder" 2>&1 | tail -8 | tee gpurun_out/r8_pytest.log
timeout 300 python scripts/bench_kernels.py --what attn --precs 1 2>&1 | tee gpurun_out/r8_kern.log
