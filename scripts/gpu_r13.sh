set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r13_pytest.log
for pdl in 1 0 1; do
MVN_PDL=$pdl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r13_bench_pdl$pdl.json 2> gpurun_out/r13_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r13_bench_pdl$pdl.json'))
print('PDL=$pdl', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['kernel_breakdown_ms'], d['loss_last'])
PY
done
tail -5 gpurun_out/r13_bench.err
