cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q -k "test_tc_clip_loss or attention" 2>&1 | tail -8 | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r08_pytest.log 2>&1; tail -12 gpurun_out/r08_pytest.log | cut -c1-300
python scripts/bench_fused.py attn 2>&1 | tail -2
for wl in c4 c3; do
timeout 400 python bench.py --steps 10 --warmup 3 --precision fused --workload $wl --no-cpu-baseline --no-sweep > gpurun_out/r08_bench_$wl.json 2> gpurun_out/r08_bench_$wl.err; tail -2 gpurun_out/r08_bench_$wl.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r08_bench_$wl.json')); print('$wl', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['kernel_breakdown_ms'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r08_c3_launches.csv python bench.py --steps 1 --warmup 3 --precision fused --workload c3 --no-cpu-baseline --no-sweep --no-graph > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r08_c3_launches.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=='ID'][0]
H=rows[hdr]; kn=H.index('Kernel Name'); mv=H.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hdr+1:]:
    try: v=float(r[mv].replace(',',''))
    except: continue
    n=r[kn].split('(')[0].replace('void ','').replace('mvn::','').replace('<unnamed>::','')[:50]
    agg[n][0]+=1; agg[n][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:28]: print(f"{k:52s} {v[0]:5d} {v[1]/1e3:9.1f} us {100*v[1]/tot:5.1f}%")
PY
