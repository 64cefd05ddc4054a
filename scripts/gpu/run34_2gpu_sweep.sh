# 2 GPUs: C4 bench line (prefetching e2e leg) followed by the C5 sweep (graph replay at <= 1024 samples per GPU)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_bench_c4_2gpu_with_c5_sweep.json 2> gpurun_out/r02c_bench_c4_2gpu_with_c5_sweep.err; tail -2 gpurun_out/r02c_bench_c4_2gpu_with_c5_sweep.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02c_bench_c4_2gpu_with_c5_sweep.json')); print('c4 2gpu', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
for r in d['c5_sweep']['sizes']: print(r.get('global_batch'), r.get('per_gpu_batch'), r.get('launch'), round(r.get('ms_per_step',0),2), round(r.get('samples_per_s',0)), round(r.get('loss_kernels_ms',0),2), r.get('error'))"
