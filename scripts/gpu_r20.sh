set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r20_pytest.log
for conc in 0 1; do for pdl in 0 1; do
MVN_CONCURRENT=$conc MVN_PDL=$pdl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r20_bench_c${conc}_p${pdl}.json 2>> gpurun_out/r20_bench.err
done; done
tail -3 gpurun_out/r20_bench.err
python - <<'PY'
import json
for c in (0,1):
  for p in (0,1):
    try:
        d=json.load(open(f'gpurun_out/r20_bench_c{c}_p{p}.json'))
        print('conc',c,'pdl',p, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), round(d['roofline']['frac'],3), {k:v['ms'] for k,v in d['kernel_breakdown_ms'].items()}, d['loss_last'])
    except Exception as e: print(c,p,'ERR',e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_mma -s 4 -c 2 -o gpurun_out/r20_attn_fwd python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r20_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_mma -s 4 -c 1 -o gpurun_out/r20_attn_bwd_a python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r20_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_mma -s 13 -c 1 -o gpurun_out/r20_attn_bwd_b python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r20_ncu4.log 2>&1
rm -f gpurun_out/*.err
ls -la gpurun_out/; du -sh gpurun_out
