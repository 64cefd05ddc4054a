cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_parity.py fp32 2>&1 | tail -6 | tee gpurun_out/r27_dp_parity.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dp_parity.py tf32 2>&1 | tail -6 | tee -a gpurun_out/r27_dp_parity.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 scripts/dp_parity.py fp32 c5 2>&1 | tail -6 | tee -a gpurun_out/r27_dp_parity.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r27_bench_c4_2gpu.json 2> gpurun_out/r27_bench_c4_2gpu.err
tail -3 gpurun_out/r27_bench_c4_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --workload c5 --steps 10 --warmup 3 > gpurun_out/r27_bench_c5_2gpu.json 2> gpurun_out/r27_bench_c5_2gpu.err
tail -3 gpurun_out/r27_bench_c5_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r27_ref_2gpu.json 2> gpurun_out/r27_ref_2gpu.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r27_*.json')):
    try:
        d=json.load(open(f)); print(f, d['n_gpus'], round(d['value']), d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
