cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r26_pytest.log 2>&1; tail -3 gpurun_out/r26_pytest.log | cut -c1-300
for wl in c4 c2; do
timeout 400 python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline --no-sweep > gpurun_out/r26_bench_$wl.json 2> gpurun_out/r26_bench.err; tail -2 gpurun_out/r26_bench.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r26_bench_$wl.json')); print('$wl', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), {k:(v['ms'],v['launches']) for k,v in d['kernel_breakdown_ms'].items()})"
done
