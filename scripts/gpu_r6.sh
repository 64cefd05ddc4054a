set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_parity.py fp32 2>&1 | tail -5 | tee gpurun_out/r6_dp_parity.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dp_parity.py tf32 2>&1 | tail -5 | tee -a gpurun_out/r6_dp_parity.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r6_bench_2gpu.json 2> gpurun_out/r6_bench_2gpu.err
tail -3 gpurun_out/r6_bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r6_bench_2gpu.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['n_gpus'])"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench_1gpu.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r6_bench_1gpu.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['n_gpus'])"
