# 8 GPUs: the scaling point the driver also measures, with the C5 sweep up to a global batch of 65 536, and DP parity at 2/4/8 ranks
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_c4_8gpu.json 2> gpurun_out/r02b_bench_c4_8gpu.err; tail -2 gpurun_out/r02b_bench_c4_8gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02b_bench_c4_8gpu.json')); print('8gpu', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['kernel_breakdown_ms']['loss']); print(json.dumps([(s['global_batch'], round(s['ms_per_step'],2), round(s['samples_per_s']), round(s['loss_kernels_ms'],2)) for s in d['c5_sweep']['sizes']]))"
timeout 110 python -m pytest tests/test_gpu_fullmodel.py -q -x -k "data_parallel and fused" 2>&1 | tail -3 | cut -c1-300
