"""Executed-instruction mix of an .ncu-rep (source page, SASS). usage: ncu_instmix.py rep [listing]"""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
cnt, tot = collections.Counter(), 0
for r in rows[2:]:
    n = int(r[ci["Instructions Executed"]])
    op = r[ci["Source"]].split()
    o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
    cnt[o] += n
    tot += n
print("warp instructions executed:", tot)
for k, v in cnt.most_common(24):
    print(f"{k:10s} {v:12d} {v / tot * 100:5.1f}%")
if len(sys.argv) > 2:
    for i, r in enumerate(rows[2:]):
        print(f"{i:4d} {int(r[ci['Instructions Executed']]):9d} {float(r[ci['Avg. Threads Executed']]):5.1f} {r[ci['Source']].strip()[:90]}")
