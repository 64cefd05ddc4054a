# first GPU pass of round 2: fused FFN parity, the full GPU suite, tf32 vs fused bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
timeout 300 python -m pytest tests/test_gpu_fused.py -x -q -s > gpurun_out/r01_fused.log 2>&1; tail -15 gpurun_out/r01_fused.log | cut -c1-300
timeout 400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fused.py > gpurun_out/r01_pytest.log 2>&1; tail -5 gpurun_out/r01_pytest.log | cut -c1-300
for prec in tf32 fused; do
timeout 200 python bench.py --steps 10 --warmup 3 --precision $prec --no-cpu-baseline > gpurun_out/r01_bench_$prec.json 2> gpurun_out/r01_bench_$prec.err; tail -2 gpurun_out/r01_bench_$prec.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r01_bench_$prec.json')); print('$prec', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['kernel_breakdown_ms'], d['loss_last'])"
done
