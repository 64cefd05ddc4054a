cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for wl in c2 c3; do
for v in 1 0 1 0; do
MVN_PDL_NEW=$v timeout 400 python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline --no-sweep > gpurun_out/r27_bench_${wl}_$v.json 2> gpurun_out/r27_bench.err; tail -1 gpurun_out/r27_bench.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/r27_bench_${wl}_$v.json')); print('$wl pdl_new=$v', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
done
done
for v in 1 0; do echo PDL_NEW=$v; MVN_PDL_NEW=$v python scripts/bench_fused.py ffn | tail -2; done
