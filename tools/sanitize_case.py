"""A small, path-covering workload for compute-sanitizer (memcheck / racecheck / initcheck):
ragged batch with gradients, a jitter case, prediction and online imputation.
usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth  # noqa: E402

Q, D, R = 2, 4, 2
ctx = api.Context(Q, D, R, workspace_bytes=256 << 20)
sizes = [37, 130, 64, 200]
pats = [synth.make_patient(D, n, seed=40 + k) for k, n in enumerate(sizes)]
sids = [ctx.add_series(*p) for p in pats]
thetas = synth.init_hyp_lmc_sm(Q, D, R, len(sizes), seed=2)
f, g, st = ctx.nlml_grad(sids, thetas, True)
assert (st == 0).all() and np.isfinite(g).all()
mean, var, st = ctx.predict(sids[:2], thetas[:2], [0, 2, 5], np.array([0, 1, 2, 3, 0], dtype=np.int32),
                            np.array([3.0, 50.0, 7.5, 100.0, 200.0], dtype=np.float32))
assert (st == 0).all()
osid = ctx.add_series(pats[1][0], np.round(pats[1][1], 0), pats[1][2], order=api.ORDER_TIME)
m, v, st = ctx.predict_online([osid], thetas[1:2])
assert st[0] == 0 and np.isfinite(m[0]).all()
os.environ["MEDGP_RL"] = "1"
ctx2 = api.Context(Q, D, R, workspace_bytes=256 << 20)
s2 = ctx2.add_series(*synth.make_patient(D, 330, seed=9))
f2, g2, st2 = ctx2.nlml_grad([s2], thetas[:1], True)
assert st2[0] == 0
print("sanitize_case ok", float(f[0]), float(f2[0]))
