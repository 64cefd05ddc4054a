set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r5_pytest.log
tail -5 gpurun_out/r5_pytest.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r5_clocks.csv &
SMI=$!
timeout 600 python bench.py > gpurun_out/r5_bench_tf32.json 2> gpurun_out/r5_bench.err
kill $SMI
timeout 600 python bench.py --precision fp32 --no-cpu-baseline > gpurun_out/r5_bench_fp32.json 2>> gpurun_out/r5_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r5_bench_ref.json 2>> gpurun_out/r5_bench.err
tail -3 gpurun_out/r5_bench.err
# launch list of one step (4th step = the per-class breakdown step): every kernel with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mvn -s 1056 -c 352 --csv --log-file gpurun_out/r5_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r5_ncu1.log 2>&1
# memory/SOL sections for the same step
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none -k regex:mvn -s 1056 -c 352 -o gpurun_out/r5_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r5_ncu2.log 2>&1
# full set + source for the top kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm|tc_wgrad|attn_.*mma' -s 1200 -c 24 -o gpurun_out/r5_top python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r5_ncu3.log 2>&1
ls -la gpurun_out/
python - <<'PY'
import json
for f in ('r5_bench_tf32','r5_bench_fp32','r5_bench_ref'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, d.get('value'), d.get('ms_per_step'), d.get('e2e'), d.get('roofline'), d.get('kernel_breakdown_ms'), d.get('cpu_baseline'))
    except Exception as e:
        print(f, 'ERR', e)
PY
