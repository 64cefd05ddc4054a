cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for comp in 0 3.52e-4; do
echo "== MVN_TRUNC_COMP=$comp"
MVN_TRUNC_COMP=$comp timeout 600 python -m pytest tests/test_gpu_fullmodel.py -q -s -k "reduced_precision" 2>&1 | grep -E "loss rel|passed|failed" | cut -c1-200
MVN_TRUNC_COMP=$comp timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fused.py -q -s -k "seq_encoder" 2>&1 | grep -E "fwd relerr|passed|failed" | cut -c1-200
done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r05_pytest.log 2>&1; tail -8 gpurun_out/r05_pytest.log | cut -c1-300
timeout 400 python bench.py --steps 5 --warmup 3 --precision fused > gpurun_out/r05_bench_fused.json 2> gpurun_out/r05_bench_fused.err; tail -2 gpurun_out/r05_bench_fused.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r05_bench_fused.json')); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['roofline']['kernel_class'], round(d['roofline']['frac'],3), d['cpu_baseline'], d.get('reference_eager_b200'))"
