set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r19_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r19_bench_tf32.json 2> gpurun_out/r19_bench.err
tail -3 gpurun_out/r19_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r19_bench_tf32.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['kernel_breakdown_ms'], d['loss_last'], d['cpu_baseline'])
PY
timeout 300 python scripts/bench_kernels.py --what fixed,gemm --precs 1 2>&1 | grep -v nobias | tee gpurun_out/r19_kern.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:_kernel -s 1060 -c 420 --csv --log-file gpurun_out/r19_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r19_ncu1.log 2>&1
wc -l gpurun_out/r19_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_mma -s 3 -c 4 -o gpurun_out/r19_attn_fwd python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r19_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_mma -s 3 -c 12 -o gpurun_out/r19_attn_bwd python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r19_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm|tc_wgrad|ln_bwd' -s 60 -c 40 -o gpurun_out/r19_tc python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r19_ncu4.log 2>&1
ls -la gpurun_out/
