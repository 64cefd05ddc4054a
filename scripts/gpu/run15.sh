cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q -k "attention or seq_encoder" 2>&1 | tail -6 | cut -c1-300
python scripts/bench_fused.py attn 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r15_pytest.log 2>&1; tail -6 gpurun_out/r15_pytest.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sweep > gpurun_out/r15_bench_c4.json 2> gpurun_out/r15_bench_c4.err; tail -2 gpurun_out/r15_bench_c4.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r15_bench_c4.json')); print('c4', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), {k:v['ms'] for k,v in d['kernel_breakdown_ms'].items()})"
