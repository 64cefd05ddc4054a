cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 28 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r47_bench_c4.json 2> gpurun_out/r47_bench_c4.err; echo "exit $?"; tail -1 gpurun_out/r47_bench_c4.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/r47_bench_c4.json')); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['config']['launch'][:40])"
