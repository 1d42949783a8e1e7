"""aggregate an ncu launch list (gpu__time_duration.sum) by kernel; optional per-launch dump"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
L = []
for r in rows:
    if len(r) > 5 and r[0] == 'ID': hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d['Metric Name'] != 'gpu__time_duration.sum': continue
        k = re.sub(r'\(.*', '', d['Kernel Name']).replace('<unnamed>::', '').replace('void ', '')
        v = float(d['Metric Value'].replace(',', ''))
        u = d['Metric Unit']
        if u == 'ns': v /= 1e3
        elif u == 'ms': v *= 1e3
        L.append((int(d['ID']), k, v, d['Grid Size'], d['Block Size']))
agg = collections.defaultdict(lambda: [0, 0.0])
for i, k, v, g, b in L:
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:70]:70s} {v[0]:4d} {v[1]:10.1f} us {v[1]/tot:.3f}")
print("total", tot)
if len(sys.argv) > 2:
    for i, k, v, g, b in L:
        if v >= float(sys.argv[2]) and 'gemm_kernel' not in k: print(i, k[:50], f"{v:.1f}", g, b)
