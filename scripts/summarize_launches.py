"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: summarize_launches.py launches.csv [last_n_launches | sweeps:FIRST:COUNT]
(sweeps:FIRST:COUNT = the launches of COUNT sweeps starting with the FIRST-th sweep (0-based) of the
process: a sweep starts with k_clust_counts when the cluster prior has hyper-parameters, else with k_prep)"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
if len(sys.argv) > 2 and sys.argv[2].startswith("sweeps:"):
    first, count = (int(v) for v in sys.argv[2].split(":")[1:3])
    marker = "k_clust_counts" if any("k_clust_counts" in r['Kernel Name'] for r in rows) else "k_prep"
    starts = [i for i, r in enumerate(rows) if marker in r['Kernel Name']]
    rows = rows[starts[first]:starts[first + count]]
elif len(sys.argv) > 2:
    rows = rows[-int(sys.argv[2]):]
agg = collections.OrderedDict()
for row in rows:
    k = re.sub(r'\(.*', '', row['Kernel Name'])
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v *= {'ns': 1, 'nsecond': 1, 'us': 1e3, 'usecond': 1e3, 'ms': 1e6, 'msecond': 1e6}.get(u, 1)
    a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print(f"launches {len(rows)}  total {tot/1e6:.3f} ms (cold-cache, serialised: compare shares)")
print(f"{'kernel':42s} {'n':>6s} {'total ms':>10s} {'avg us':>10s} {'max us':>10s} {'share':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:42s} {a[0]:6d} {a[1]/1e6:10.3f} {a[1]/a[0]/1e3:10.2f} {a[2]/1e3:10.2f} {a[1]/tot*100:6.1f}%")
