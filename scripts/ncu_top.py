"""Top stall locations of one kernel from `ncu -i X.ncu-rep --page source --csv` (development aid)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]
si, src, ex = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
stall_cols = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
data = [r for r in rows[2:] if len(r) > max(stall_cols) and r[si].isdigit()]
tot = sum(int(r[si]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for c in stall_cols:
        agg[h[c]] = agg.get(h[c], 0) + int(r[c] or 0)
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
top = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:topn]
for i in sorted(top):
    r = data[i]
    st = sorted([(int(r[c]), h[c][6:]) for c in stall_cols if int(r[c] or 0) > 0], reverse=True)[:2]
    print(f"{i:5d} {int(r[si]):6d} {100*int(r[si])/tot:5.1f}% ex={r[ex]:>8s} {r[src].strip()[:64]:64s} {st}")
