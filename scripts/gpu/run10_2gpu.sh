# 2 GPUs: NCCL data-parallel parity (incl. overlapped gradient buckets, SyncBN, classifier head) and the 2-GPU bench with the C5 sweep
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tail -2
timeout 900 python -m pytest tests/test_gpu_fullmodel.py -q -k "data_parallel" 2>&1 | tail -6 | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r10_bench_2gpu.json 2> gpurun_out/r10_bench_2gpu.err; tail -3 gpurun_out/r10_bench_2gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r10_bench_2gpu.json')); print('2gpu', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['kernel_breakdown_ms']['loss']); print(json.dumps(d.get('c5_sweep'), indent=0)[:1500])"
