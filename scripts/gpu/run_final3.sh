# Final measurement pass of the round's third session on one B200: GPU test suite, smoke, bench lines for every workload (C4 with the CPU
# baseline and the C5 sweep, as the driver runs it), the reference arm.  Kernels are unchanged since run_final2*.sh: the ncu captures
# under profiles/r02b_* still describe this HEAD's kernels.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=r02c_final
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv | tail -1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; tail -2 gpurun_out/${T}_bench_c4.err | cut -c1-300
for wl in c3 c5 c2 c1; do
timeout 400 python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline --no-sweep > gpurun_out/${T}_bench_$wl.json 2> gpurun_out/${T}_bench_$wl.err; tail -2 gpurun_out/${T}_bench_$wl.err | cut -c1-300
done
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_reference_arm.json 2> gpurun_out/${T}_reference_arm.err; tail -2 gpurun_out/${T}_reference_arm.err | cut -c1-300
for f in c4 c3 c5 c2 c1; do python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_$f.json')); r=d['roofline']; print('$f', round(d['value']), 'samples/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), '; top', r['kernel_class'], round(r['frac'],4), '; step frac padded', round(r['step_frac_of_tensor_peak']['padded'],4), '; GB/step', round(d['bytes_per_step']['total']/1e9,2))"; done
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_c4.json')); print(d.get('cpu_baseline')); print(d.get('reference_eager_b200')); print(json.dumps(d.get('c5_sweep'))[:900])
r=json.load(open('gpurun_out/${T}_reference_arm.json')); print('reference arm', round(r['value'],1), r['cpu_baseline']['sample'][:160])"
