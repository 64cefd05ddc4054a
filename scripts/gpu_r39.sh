cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 100 python scripts/tc_diag.py 2>&1 | grep -v "kernel added\|bad rows" | cut -c1-200 | tee gpurun_out/r39_diag.txt
timeout 150 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/r39_pytest.log 2>&1; tail -3 gpurun_out/r39_pytest.log | cut -c1-200
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r39_bench_c4.json 2> gpurun_out/r39_bench_c4.err; tail -2 gpurun_out/r39_bench_c4.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r39_bench_c4.json')); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['kernel_breakdown_ms'], d['roofline']['frac'])
PY
