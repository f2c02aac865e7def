"""Stage timeline of the multi-stream step (diagnostics): runs a few C2 steps with
MEDGP_TIMELINE, then prints per-stream stage intervals of the last call and how many streams
are in which stage over time.  usage: python tools/timeline.py [patients] [n]"""
import collections
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = os.path.join(tempfile.mkdtemp(), "timeline.txt")
os.environ["MEDGP_TIMELINE"] = path
from medgp_b200 import api, synth  # noqa: E402

patients = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
Q, D, R = 5, 24, 8
ctx = api.Context(Q, D, R, workspace_bytes=16 << 30)
sids = [ctx.add_series(*synth.make_patient(D, n, seed=i)) for i in range(patients)]
thetas = synth.init_hyp_lmc_sm(Q, D, R, patients, seed=718)
ctx.profile(True)
for _ in range(3):
    ctx.nlml_grad(sids, thetas, True)
calls = open(path).read().split("# end of call\n")
rows = [tuple(float(v) for v in l.split()) for l in calls[-2].splitlines() if l.strip()]
by_stream = collections.defaultdict(list)
for s, st, a, b in rows:
    by_stream[int(s)].append((a, b, api.STAGES[int(st)]))
t_end = max(b for _, _, _, b in rows)
print(f"step: {t_end:.3f} ms over {len(by_stream)} streams (event timing on: launches are not graph-replayed)")
for s in sorted(by_stream):
    agg = collections.OrderedDict()
    for a, b, name in sorted(by_stream[s]):
        if name in agg:
            agg[name] = (agg[name][0], b, agg[name][2] + (b - a))
        else:
            agg[name] = (a, b, b - a)
    print(f"stream {s}: " + "  ".join(f"{k}[{v[0]:.2f}-{v[1]:.2f} busy {v[2]:.2f}]" for k, v in agg.items()))
