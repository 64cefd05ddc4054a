cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/fin2_pytest.log 2>&1; tail -3 gpurun_out/fin2_pytest.log | cut -c1-200
timeout 150 python bench.py --steps 10 --warmup 3 > gpurun_out/fin2_bench_c4.json 2> gpurun_out/fin2_bench_c4.err; tail -1 gpurun_out/fin2_bench_c4.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/fin2_bench_c4.json')); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['roofline']['frac'], d['cpu_baseline'], d['gpu_launches'], d['clocks'])"
