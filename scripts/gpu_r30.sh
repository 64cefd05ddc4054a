cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r30_bench_c4_2gpu.json 2> gpurun_out/r30_bench_c4_2gpu.err
tail -5 gpurun_out/r30_bench_c4_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --workload c5 --steps 10 --warmup 3 > gpurun_out/r30_bench_c5_2gpu.json 2> gpurun_out/r30_bench_c5_2gpu.err
tail -5 gpurun_out/r30_bench_c5_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --no-graph --steps 10 --warmup 3 > gpurun_out/r30_bench_c4_2gpu_nograph.json 2> gpurun_out/r30_bench_c4_2gpu_nograph.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r30_*.json')):
    try:
        d=json.load(open(f)); print(f, d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['step_ms_rank0'])
    except Exception as e: print(f, 'ERR', e)
PY
wc -l gpurun_out/r30_*.json
