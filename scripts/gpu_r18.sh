set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r18_pytest.log
timeout 300 python scripts/bench_kernels.py --what fixed,gemm --precs 1 2>&1 | grep -v nobias | tee gpurun_out/r18_kern.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r18_bench_tf32.json 2> gpurun_out/r18_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r18_bench_tf32.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['kernel_breakdown_ms'], d['loss_last'])
PY
