# C1 (configs/maven-lite.yaml as shipped) full-model parity + a bench line; prefetch tests with pinned dropout seeds.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=r02c
timeout 500 python -m pytest tests/test_gpu_graph.py tests/test_gpu_fullmodel.py -m gpu -q -s -k "prefetch or c1" > gpurun_out/${T}_pytest_c1.log 2>&1; grep -h "c1 \|passed\|failed\|Error" gpurun_out/${T}_pytest_c1.log | cut -c1-300 | tail -12
timeout 400 python bench.py --steps 10 --warmup 3 --workload c1 --no-cpu-baseline --no-sweep > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err; tail -3 gpurun_out/${T}_bench_c1.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_c1.json')); r=d['roofline']; print('c1: value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'top', r['kernel_class'], round(r['frac'],4), r['step_frac_of_tensor_peak'], d['kernel_breakdown_ms'])"
