"""Per-kernel totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list: python scripts/launch_summary.py file.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]; k = h.index("Kernel Name"); v = h.index("Metric Value")
d = collections.OrderedDict()
for r in rows[1:]:
    n = r[k].split("(")[0][:90]
    d.setdefault(n, [0, 0.0]); d[n][0] += 1; d[n][1] += float(r[v].replace(",", "")) / 1e6
tot = sum(x[1] for x in d.values())
for n, (c, ms) in sorted(d.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:9.3f} ms {c:4d}  {100 * ms / tot:5.1f}%  {n}")
print(f"TOTAL {tot:.3f} ms, {sum(x[0] for x in d.values())} launches")
