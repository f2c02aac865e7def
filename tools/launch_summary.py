"""Summarise an ncu gpu__time_duration launch list (csv): python tools/launch_summary.py file [skip_fraction]"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = rows[int(len(rows) * skip):]
agg = collections.OrderedDict()
for row in rows:
    k = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    agg.setdefault(k, []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:22s} n={len(v):3d} total={sum(v):9.1f}us share={sum(v)/tot*100:5.1f}%  each=", " ".join(f"{x:.0f}" for x in v[:20]))
print(f"total {tot:.1f} us")
