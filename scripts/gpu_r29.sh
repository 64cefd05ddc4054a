cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graph.py -x -q > gpurun_out/r29_pytest.log 2>&1; tail -25 gpurun_out/r29_pytest.log
for w in c4 c3 c5; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r29_bench_$w.json 2> gpurun_out/r29_bench_$w.err; tail -3 gpurun_out/r29_bench_$w.err
done
timeout 300 python bench.py --workload c4 --no-graph --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r29_bench_c4_nograph.json 2> gpurun_out/r29_bench_c4_nograph.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r29_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'], d['step_ms_rank0'][:4])
    except Exception as e: print(f, 'ERR', e)
PY
