set -x
cd $GRAFT_REPO_ROOT
timeout 300 python scripts/bench_kernels.py --what gemm --precs 1 2>&1 | grep wgrad | tee gpurun_out/r11_kern.log
