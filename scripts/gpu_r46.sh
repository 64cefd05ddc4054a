cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 60 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r46_bench_c3.json 2> gpurun_out/r46_bench_c3.err; tail -2 gpurun_out/r46_bench_c3.err | cut -c1-250
python -c "
import json; d=json.load(open('gpurun_out/r46_bench_c3.json')); print(round(d['value']), round(d['ms_per_step'],3), d['e2e'], d['loss_last'])"
