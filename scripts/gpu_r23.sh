set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r23_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r23_pytest.log
tail -5 gpurun_out/r23_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r23_bench_c4.json 2> gpurun_out/r23_bench_c4.err; tail -2 gpurun_out/r23_bench_c4.err
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r23_bench_c3.json 2> gpurun_out/r23_bench_c3.err; tail -2 gpurun_out/r23_bench_c3.err
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r23_bench_c5.json 2> gpurun_out/r23_bench_c5.err; tail -2 gpurun_out/r23_bench_c5.err
timeout 300 python bench.py --workload c5 --batch 4096 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r23_bench_c5_b4096.json 2> gpurun_out/r23_bench_c5_b4096.err; tail -2 gpurun_out/r23_bench_c5_b4096.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r23_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['kernel_breakdown_ms'])
    except Exception as e: print(f, 'ERR', e)
PY
