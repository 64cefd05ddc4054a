cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r04_pytest.log 2>&1; tail -15 gpurun_out/r04_pytest.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 400 python bench.py --steps 5 --warmup 3 --precision fused > gpurun_out/r04_bench_fused.json 2> gpurun_out/r04_bench_fused.err; tail -2 gpurun_out/r04_bench_fused.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r04_bench_fused.json')); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['roofline']['kernel_class'], round(d['roofline']['frac'],3), d['cpu_baseline'], d.get('reference_eager_b200'))"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r04_bench_ref.json 2> gpurun_out/r04_bench_ref.err; tail -2 gpurun_out/r04_bench_ref.err | cut -c1-300; cut -c1-600 gpurun_out/r04_bench_ref.json
nproc
