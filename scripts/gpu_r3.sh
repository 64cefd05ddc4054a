set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -s -k "attention or seq_encoder" 2>&1 | tail -25 > gpurun_out/r3_pytest.log
tail -25 gpurun_out/r3_pytest.log
timeout 300 python scripts/bench_kernels.py --what attn > gpurun_out/r3_kern.log 2>&1
cat gpurun_out/r3_kern.log
timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/r3_bench_tf32.json 2> gpurun_out/r3_bench.err
cat gpurun_out/r3_bench_tf32.json; tail -5 gpurun_out/r3_bench.err
