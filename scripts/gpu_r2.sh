set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q  2>&1 | tail -25 > gpurun_out/r2_pytest.log
tail -25 gpurun_out/r2_pytest.log
timeout 300 python scripts/bench_kernels.py --what gemm > gpurun_out/r2_kern.log 2>&1
cat gpurun_out/r2_kern.log
