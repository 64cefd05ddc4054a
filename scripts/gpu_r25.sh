set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
scripts/micro/mma_rate > gpurun_out/r25_mma_rate.txt 2>&1; cat gpurun_out/r25_mma_rate.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "conv or training_step" > gpurun_out/r25_pytest.log 2>&1; tail -3 gpurun_out/r25_pytest.log
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r25_bench_c3.json 2> gpurun_out/r25_bench_c3.err; tail -2 gpurun_out/r25_bench_c3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r25_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['kernel_breakdown_ms'])
    except Exception as e: print(f, 'ERR', e)
PY
