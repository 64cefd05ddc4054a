set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -s 2>&1 | tail -25 > gpurun_out/r4_pytest.log
tail -25 gpurun_out/r4_pytest.log
timeout 300 python scripts/bench_kernels.py --what gemm --precs 1 > gpurun_out/r4_kern.log 2>&1
cat gpurun_out/r4_kern.log
timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/r4_bench_tf32.json 2> gpurun_out/r4_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r4_bench_tf32.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['kernel_breakdown_ms'])
PY
tail -5 gpurun_out/r4_bench.err
