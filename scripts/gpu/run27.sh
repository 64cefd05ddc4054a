cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for wl in c4 c2 c3 c5; do
timeout 400 python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline --no-sweep > gpurun_out/r27_bench_${wl}.json 2> gpurun_out/r27_bench.err; tail -1 gpurun_out/r27_bench.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/r27_bench_${wl}.json')); print('$wl', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
done
