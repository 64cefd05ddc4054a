cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r26_pytest.log 2>&1; tail -3 gpurun_out/r26_pytest.log | cut -c1-300
for pdl in 1 0; do
MVN_PDL=$pdl timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sweep > gpurun_out/r26_bench_c4_$pdl.json 2> gpurun_out/r26_bench_c4.err; tail -2 gpurun_out/r26_bench_c4.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r26_bench_c4_$pdl.json')); print('c4 pdl=$pdl', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), {k:(v['ms'],v['launches']) for k,v in d['kernel_breakdown_ms'].items()})"
done
