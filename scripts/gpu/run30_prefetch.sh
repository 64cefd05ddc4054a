# Input prefetch (GraphedTrainStep double_buffer): GPU tests, then the end-to-end leg with and without it on C4 and C3.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=r02c
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${T}_pytest_gpu.log | cut -c1-400
for wl in c4 c3; do
for pf in on off; do
fl=""; [ $pf = off ] && fl="--no-prefetch"
timeout 300 python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline --no-sweep $fl > gpurun_out/${T}_bench_${wl}_pf${pf}.json 2> gpurun_out/${T}_bench_${wl}_pf${pf}.err; tail -2 gpurun_out/${T}_bench_${wl}_pf${pf}.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_${wl}_pf${pf}.json')); print('$wl prefetch $pf: value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'loss', d.get('loss_last'), d['config'].get('launch','')[:80])"
done; done
