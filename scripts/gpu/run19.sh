cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "conv or clip3 or c3 or c5 or training_step or fullmodel or graph" 2>&1 | tail -8 | cut -c1-400
for f in 1 0; do
MVN_CONV_FUSED=$f timeout 400 python bench.py --steps 20 --warmup 5 --workload c3 --no-cpu-baseline --no-sweep > gpurun_out/r19_bench_c3_$f.json 2> gpurun_out/r19_bench_c3_$f.err; tail -2 gpurun_out/r19_bench_c3_$f.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r19_bench_c3_$f.json')); print('c3 fused=$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), {k:(v['ms'],v['launches']) for k,v in d['kernel_breakdown_ms'].items()})"
done
