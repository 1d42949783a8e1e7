"""top stall-sample SASS instructions of the first kernel in `ncu --page source --csv --print-source sass` output"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name': break          # only the first kernel of the report
    if len(r) == len(hdr) and r[0] != 'Address': data.append(r)
num = lambda s: int(float(s)) if s not in ("", None) else 0
tot = sum(num(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
for i, r in enumerate(data): r.append(i)
top = sorted(data, key=lambda r: -num(r[ix['# Samples']]))[:n]
for r in top:
    stalls = {h: num(r[ix[h]]) for h in hdr if h.startswith('stall_') and 'Not Issued' not in h}
    s = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print(str(r[-1]).rjust(5), r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(9), r[ix['Source']][:80].ljust(80), s)
