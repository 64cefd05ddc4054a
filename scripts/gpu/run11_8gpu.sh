# 8 GPUs: the scaling point the driver also measures, with the C5 sweep up to a global batch of 65 536, and DP parity at 2/4/8 ranks
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tail -8 | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r11_bench_8gpu.json 2> gpurun_out/r11_bench_8gpu.err; tail -3 gpurun_out/r11_bench_8gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r11_bench_8gpu.json')); print('8gpu', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['kernel_breakdown_ms']['loss']); print(json.dumps(d.get('c5_sweep'), indent=0)[:2500])"
timeout 900 python -m pytest tests/test_gpu_fullmodel.py -q -k "data_parallel and fused" 2>&1 | tail -6 | cut -c1-400
