set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r17_pytest.log
timeout 300 python scripts/bench_kernels.py --what gemm --precs 1 2>&1 | grep -v nobias | grep wgrad | tee gpurun_out/r17_kern.log
