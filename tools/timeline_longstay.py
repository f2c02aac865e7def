"""Per-launch timeline of the factorisation of `count` copies of one long series in flight
(MEDGP_TIMELINE, profile mode: no graph replay, events around every stage interval).
usage: python tools/timeline_longstay.py [count] [n]  -> prints stream 0's intervals and a summary"""
import collections
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = os.path.join(tempfile.mkdtemp(), "timeline.txt")
os.environ["MEDGP_TIMELINE"] = path
from medgp_b200 import api, synth  # noqa: E402

count = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
Q, D, R = 5, 24, 8
meta, x, y = synth.make_patient(D, n, seed=4000, T=1200.0)
ctx = api.Context(Q, D, R, workspace_bytes=12 << 30)
sid = ctx.add_series(meta, x, y)
thetas = synth.init_hyp_lmc_sm(Q, D, R, count, seed=4)
ctx.profile(True)
for _ in range(3):
    ctx.nlml_grad([sid] * count, thetas, False)
calls = open(path).read().split("# end of call\n")
rows = [tuple(float(v) for v in l.split()) for l in calls[-2].splitlines() if l.strip()]
by_stream = collections.defaultdict(list)
for s, st, a, b in rows:
    by_stream[int(s)].append((a, b, api.STAGES[int(st)]))
t_end = max(b for _, _, _, b in rows)
print(f"call: {t_end:.3f} ms over {len(by_stream)} streams")
for s in sorted(by_stream):
    iv = sorted(by_stream[s])
    busy = collections.Counter()
    for a, b, name in iv:
        busy[name] += b - a
    gaps = sum(max(0.0, iv[k + 1][0] - iv[k][1]) for k in range(len(iv) - 1))
    print(f"stream {s}: first {iv[0][0]:.3f} last {iv[-1][1]:.3f} | " + " ".join(f"{k}={v:.3f}" for k, v in busy.items()) + f" | gaps {gaps:.3f} ms")
iv = sorted(by_stream[min(by_stream)])
print("stream", min(by_stream), "intervals (start, duration us, stage), first 60 and a middle slice:")
for a, b, name in iv[:60] + iv[len(iv) // 2: len(iv) // 2 + 30]:
    print(f"  {a * 1e3:9.1f} {1e3 * (b - a):8.1f} {name}")
