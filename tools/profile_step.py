"""One or a few steps of the bench workload (for ncu launch lists / --set full captures).
usage: python tools/profile_step.py [steps] [patients] [n_points] [want_grad]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
patients = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = int(sys.argv[3]) if len(sys.argv) > 3 else 500
want_grad = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
Q, D, R = 5, 24, 8
ctx = api.Context(Q, D, R, workspace_bytes=16 << 30)
sids = [ctx.add_series(*synth.make_patient(D, n, seed=i)) for i in range(patients)]
thetas = synth.init_hyp_lmc_sm(Q, D, R, patients, seed=718)
ctx.profile(True)
for _ in range(steps):
    f, g, st = ctx.nlml_grad(sids, thetas, want_grad)
t = ctx.stage_times()
print({k: (round(v["ms"] / steps, 4), v["launches"] // steps) for k, v in t.items() if k != "evals"})
assert (st == 0).all()
