# C5 sweep with graph replay at the launch-bound sizes: 1 GPU, sizes 1024 (replay) and 2048 (eager)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sweep-max-per-gpu 2048 > gpurun_out/r02c_bench_c4_sweep1.json 2> gpurun_out/r02c_bench_c4_sweep1.err; tail -3 gpurun_out/r02c_bench_c4_sweep1.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02c_bench_c4_sweep1.json')); print(round(d['value']), json.dumps(d.get('c5_sweep'))[:1200])"
