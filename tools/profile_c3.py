"""A reduced C3 cohort (same size distribution and model as bench.py: n ~ U{300..1500}, D=24, Q=5, R=8,
5 theta per patient) for ncu launch lists and --set full captures: a full C3 step is 20480 evaluations,
far too long under a profiler.  Prints the per-stage times and the sum of n^2 / n^3 of one step.
usage: python tools/profile_c3.py [steps] [patients] [inits]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from medgp_b200 import api, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
patients = int(sys.argv[2]) if len(sys.argv) > 2 else 128
inits = int(sys.argv[3]) if len(sys.argv) > 3 else 5
Q, D, R = bench.Q, bench.D, bench.R
sizes = bench.cohort_sizes()[:patients]
ctx = api.Context(Q, D, R, workspace_bytes=24 << 30)
sids = np.repeat([ctx.add_series(*bench.cohort_patient(i, int(n))) for i, n in enumerate(sizes)], inits)
thetas = synth.init_hyp_lmc_sm(Q, D, R, bench.THETA_POOL, seed=718)[np.arange(len(sids)) % bench.THETA_POOL]
ctx.profile(True)
for _ in range(steps):
    f, g, st = ctx.nlml_grad(sids, thetas, True)
t = ctx.stage_times()
assert (st == 0).all() or os.environ.get('AB_NOASSERT')
print(json.dumps({"patients": patients, "inits": inits, "evals_per_step": len(sids),
                  "sum_n2_per_step": float((sizes.astype(float) ** 2).sum() * inits),
                  "sum_n3_per_step": float((sizes.astype(float) ** 3).sum() * inits),
                  "stage_ms_per_step": {k: v["ms"] / steps for k, v in t.items() if k != "evals"},
                  "stage_launches_per_step": {k: v["launches"] // steps for k, v in t.items() if k != "evals"}}))
