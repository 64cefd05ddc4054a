cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/r35_pytest.log 2>&1; tail -4 gpurun_out/r35_pytest.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r35_bench_c4.json 2> gpurun_out/r35_bench_c4.err; tail -2 gpurun_out/r35_bench_c4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r35_bench_c4.json')); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['kernel_breakdown_ms'], d['roofline']['frac'])
PY
