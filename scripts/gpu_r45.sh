cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_graph.py -x -q -k "attention or seq_encoder or training_step or graphed" > gpurun_out/r45_pytest.log 2>&1; tail -4 gpurun_out/r45_pytest.log | cut -c1-200
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r45_bench_c4.json 2> gpurun_out/r45_bench_c4.err; tail -1 gpurun_out/r45_bench_c4.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/r45_bench_c4.json')); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['kernel_breakdown_ms']['attn_fwd'], d['kernel_breakdown_ms']['attn_bwd'])"
