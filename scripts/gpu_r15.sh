set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r15_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r15_bench_tf32.json 2> gpurun_out/r15_bench.err
tail -3 gpurun_out/r15_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r15_bench_tf32.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['kernel_breakdown_ms'], d['loss_last'], d['cpu_baseline'])
PY
timeout 300 python scripts/bench_kernels.py --what fixed,gemm --precs 1 2>&1 | grep -v nobias | tee gpurun_out/r15_kern.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:_kernel -s 1060 -c 420 --csv --log-file gpurun_out/r15_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r15_ncu1.log 2>&1
wc -l gpurun_out/r15_launches.csv
