cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_graph.py -x -q > gpurun_out/r43_pytest.log 2>&1; tail -4 gpurun_out/r43_pytest.log | cut -c1-200
timeout 100 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/fin_bench_c2.json 2> gpurun_out/fin_bench_c2.err; tail -1 gpurun_out/fin_bench_c2.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/fin_bench_c2.json')); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
