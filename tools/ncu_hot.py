"""Top stalled SASS instructions of an .ncu-rep (source page). usage: ncu_hot.py rep [min_pct]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = sum(int(r[ci["# Samples"]]) for r in data)
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
print("total samples", tot)
agg = {h: sum(int(r[ci[h]]) for r in data) for h in stall_cols}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda t: -t[1])[:8]})
for i, r in enumerate(data):
    s = int(r[ci["# Samples"]])
    if s >= tot * minpct / 100:
        top = sorted(((int(r[ci[h]]), h) for h in stall_cols), reverse=True)[:2]
        print(f"{i:4d} {s/tot*100:5.1f}%  {r[ci['Source']].strip()[:80]:80s} {top}")
