set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r21_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/r21_smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r21_bench.json 2> gpurun_out/r21_bench.err
tail -3 gpurun_out/r21_bench.err
for conc in 0 1; do
MVN_CONCURRENT=$conc timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r21_bench_c${conc}.json 2>> gpurun_out/r21_bench.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r21_ref.json 2>> gpurun_out/r21_bench.err
python - <<'PY'
import json
for f in ('r21_bench','r21_bench_c0','r21_bench_c1'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), round(d['roofline']['frac'],3), {k:v['ms'] for k,v in d['kernel_breakdown_ms'].items()}, d['loss_last'])
    except Exception as e: print(f,'ERR',e)
print(open('gpurun_out/r21_ref.json').read()[:600])
PY
timeout 300 python scripts/bench_kernels.py > gpurun_out/r21_kernels.txt 2>&1; tail -50 gpurun_out/r21_kernels.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r21_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r21_ncu1.log 2>&1
ls -la gpurun_out/; du -sh gpurun_out
