"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals of the LAST step.
    python tools/launch_summary.py file.csv launches_per_step"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[h]; ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit')
data = [r for r in rows[h + 1:] if len(r) > vi]
per = int(sys.argv[2]) if len(sys.argv) > 2 else len(data)
last = data[-per:]
tot = collections.Counter(); cnt = collections.Counter()
for r in last:
    v = float(r[vi].replace(',', ''))
    if r[ui] in ('nsecond', 'ns'): v /= 1e6
    elif r[ui] in ('usecond', 'us'): v /= 1e3
    nm = re.sub(r'^void ', '', r[ki]); nm = re.sub(r'^ub::', '', nm); nm = nm.split('(')[0]
    tot[nm] += v; cnt[nm] += 1
s = sum(tot.values())
print(f"{len(data)} launches in the file; last {per}: {s:.2f} ms serialised")
for nm, v in tot.most_common(45):
    print(f"  {v:8.3f} ms {100 * v / s:5.1f}% {cnt[nm]:4d}  {nm[:90]}")
