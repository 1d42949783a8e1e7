"""dynamic opcode histogram (warp instructions executed) + hottest SASS lines of kernel #k in `ncu --page source --csv --print-source sass` output"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
s = starts[k]; e = starts[k + 1] if k + 1 < len(starts) else len(rows)
print(rows[s][1])
hdr = rows[s + 1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
num = lambda v: int(float(v)) if v not in ("", None) else 0
tot = sum(num(r[ix['Instructions Executed']]) for r in data)
print("warp instructions executed", tot, "static", len(data))
h = collections.Counter()
for r in data:
    op = re.sub(r'^@!?U?P\w+\s+', '', r[ix['Source']].strip()).split()[0].split('.')[0]
    h[op] += num(r[ix['Instructions Executed']])
for op, c in h.most_common(22):
    print(f"{op:10s} {c / 1e6:8.2f}M {100 * c / tot:5.1f}%")
if top:
    mx = max(num(r[ix['Instructions Executed']]) for r in data)
    for i, r in enumerate(data):
        if num(r[ix['Instructions Executed']]) >= mx * 0.45:
            print(str(i).rjust(5), r[ix['Instructions Executed']].rjust(9), r[ix['# Samples']].rjust(6), r[ix['Source']][:100])
