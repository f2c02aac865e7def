"""Ragged cohort throughput (BASELINE.json config 3 shape): patients with n ~ U{300..1500},
several hyper-parameter vectors per patient, one NLML+gradient evaluation each per step.
usage: python tools/bench_c3.py [patients] [inits] [steps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth  # noqa: E402

patients = int(sys.argv[1]) if len(sys.argv) > 1 else 256
inits = int(sys.argv[2]) if len(sys.argv) > 2 else 5
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
Q, D, R = 5, 24, 8
rng = np.random.default_rng(3)
sizes = rng.integers(300, 1501, patients)
ctx = api.Context(Q, D, R)
sids = [ctx.add_series(*synth.make_patient(D, int(n), seed=1000 + i, T=240.0 * n / 500.0)) for i, n in enumerate(sizes)]
sid_b = np.repeat(sids, inits)
thetas = synth.init_hyp_lmc_sm(Q, D, R, patients * inits, seed=718)
th = ctx.pinned(thetas.shape)
th[...] = thetas
outs = (ctx.pinned((len(sid_b),)), ctx.pinned((len(sid_b), ctx.P)), ctx.pinned((len(sid_b),), np.int32))
ctx.nlml_grad(sid_b, th, True, out=outs)
t0 = time.perf_counter()
for _ in range(steps):
    f, g, st = ctx.nlml_grad(sid_b, th, True, out=outs)
dt = (time.perf_counter() - t0) / steps
flops = float(np.sum(sizes.astype(np.float64) ** 3)) * inits  # potrf + trtri + lauum = n^3
print(json.dumps({"patients": patients, "inits": inits, "evals_per_step": len(sid_b), "n_min": int(sizes.min()),
                  "n_max": int(sizes.max()), "n_mean": float(sizes.mean()), "s_per_step": dt,
                  "evals_per_s_e2e": len(sid_b) / dt, "linear_algebra_tflops_e2e": flops / dt / 1e12,
                  "failed": int((st < 0).sum()), "jittered": int((st > 0).sum())}))
