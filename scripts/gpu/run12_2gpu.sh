cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r12_bench_1gpu.json 2> gpurun_out/r12_bench_1gpu.err; tail -3 gpurun_out/r12_bench_1gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r12_bench_1gpu.json')); print('1gpu', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['kernel_breakdown_ms']); print(json.dumps(d.get('c5_sweep'), indent=0)[:1800])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r12_bench_2gpu.json 2> gpurun_out/r12_bench_2gpu.err; tail -3 gpurun_out/r12_bench_2gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r12_bench_2gpu.json')); print('2gpu', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['kernel_breakdown_ms']['loss']); print(json.dumps(d.get('c5_sweep'), indent=0)[:1800])"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "convmixer or training_step" 2>&1 | tail -3
