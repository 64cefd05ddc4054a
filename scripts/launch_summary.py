"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  Usage: python scripts/launch_summary.py file.csv [top] [filter]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
flt = sys.argv[3] if len(sys.argv) > 3 else ""
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    k = d['Kernel Name'].replace('mvn::<unnamed>::', '').replace('void ', '')[:64]
    v = float(d['Metric Value'].replace(',', '')); u = d['Metric Unit']
    v = v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
print('total us', round(tot), 'launches', sum(v[0] for v in agg.values()))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if flt and flt not in k: continue
    top -= 1
    if top < 0: break
    print(f'  {k:64s} {n:5d} {t:9.0f} us {100*t/tot:5.1f} %  {t/n:7.1f} us/launch')
