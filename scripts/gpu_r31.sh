cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
date +%T
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r31_bench_c4_2gpu.json 2> gpurun_out/r31_bench_c4_2gpu.err; echo "exit $?"
date +%T
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --workload c5 --steps 10 --warmup 3 > gpurun_out/r31_bench_c5_2gpu.json 2> gpurun_out/r31_bench_c5_2gpu.err; echo "exit $?"
date +%T
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r31_*.json')):
    try:
        d=json.load(open(f)); print(f, d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))
    except Exception as e: print(f, 'ERR', e)
PY
wc -l gpurun_out/r31_*.json
