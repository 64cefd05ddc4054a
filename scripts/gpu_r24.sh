set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r24_c3_launches.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r24_ncu.log 2>&1
for i in 1 2 3; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r24_bench_c4_$i.json 2> gpurun_out/r24_bench_c4_$i.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r24_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['step_ms_rank0'])
    except Exception as e: print(f, 'ERR', e)
PY
