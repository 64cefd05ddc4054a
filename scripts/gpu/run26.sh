cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r26_pytest.log 2>&1; tail -3 gpurun_out/r26_pytest.log | cut -c1-300
for v in 1 0 1 0; do
MVN_ATTN_ORDER=$v timeout 400 python bench.py --steps 20 --warmup 5 --workload c4 --no-cpu-baseline --no-sweep > gpurun_out/r26_bench_c4_$v.json 2> gpurun_out/r26_bench.err; tail -2 gpurun_out/r26_bench.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r26_bench_c4_$v.json')); print('c4 order=$v', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), {k:(v['ms'],v['launches']) for k,v in d['kernel_breakdown_ms'].items() if k.startswith('attn')})"
done
