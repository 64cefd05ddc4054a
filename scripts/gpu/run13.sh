cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -q -x 2>&1 | tail -12 | cut -c1-300
MVN_FFN_FWD=1 python scripts/bench_fused.py ffn 2>&1 | tail -2
MVN_FFN_FWD=3 python scripts/bench_fused.py ffn 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r13_pytest.log 2>&1; tail -6 gpurun_out/r13_pytest.log | cut -c1-300
for wl in c4 c3; do
timeout 400 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --no-sweep > gpurun_out/r13_bench_$wl.json 2> gpurun_out/r13_bench_$wl.err; tail -2 gpurun_out/r13_bench_$wl.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r13_bench_$wl.json')); print('$wl', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), {k:v['ms'] for k,v in d['kernel_breakdown_ms'].items()})"
done
