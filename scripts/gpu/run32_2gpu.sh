# 2 GPUs: data-parallel parity test + C4 bench line with the prefetching end-to-end leg (both graphs capture the NCCL all-reduce)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fullmodel.py -q -k "data_parallel" 2>&1 | tail -3 | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-sweep > gpurun_out/r02c_bench_c4_2gpu.json 2> gpurun_out/r02c_bench_c4_2gpu.err; tail -2 gpurun_out/r02c_bench_c4_2gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02c_bench_c4_2gpu.json')); print('c4 2gpu', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e'].get('input_prefetch'), d['config'].get('launch'))"
