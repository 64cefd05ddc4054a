cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
date +%T
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/fin_pytest.log 2>&1; tail -6 gpurun_out/fin_pytest.log | cut -c1-200
date +%T
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/fin_smoke.log 2>&1; tail -2 gpurun_out/fin_smoke.log | cut -c1-200
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/fin_bench_c4.json 2> gpurun_out/fin_bench_c4.err; tail -1 gpurun_out/fin_bench_c4.err | cut -c1-200
timeout 120 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/fin_bench_c3.json 2> gpurun_out/fin_bench_c3.err
timeout 120 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/fin_bench_c5.json 2> gpurun_out/fin_bench_c5.err
timeout 120 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/fin_bench_c2.json 2> gpurun_out/fin_bench_c2.err; tail -1 gpurun_out/fin_bench_c2.err | cut -c1-200
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin_reference_arm.json 2> gpurun_out/fin_reference_arm.err
date +%T
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/fin_ncu.log 2>&1; tail -2 gpurun_out/fin_ncu.log | cut -c1-200; wc -l gpurun_out/fin_launches.csv
date +%T
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/fin_*.json')):
    try:
        d=json.load(open(f)); print(f, d.get('n_gpus'), round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('roofline',{}).get('frac'), d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
