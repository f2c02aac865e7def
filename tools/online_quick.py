"""predict_online over a batch of time-ordered series, repeated calls, wall time per call (graphs vs direct launches)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth
Q, D, R = 5, 24, 8
sizes = np.random.default_rng(5).integers(200, 501, 512)
ctx = api.Context(Q, D, R)
theta = synth.init_hyp_lmc_sm(Q, D, R, 3, seed=718)[2]
sids = [ctx.add_series(*synth.make_patient(D, int(n), seed=50000 + k, T=240.0 * n / 500.0), order=api.ORDER_TIME) for k, n in enumerate(sizes)]
thetas = np.tile(theta, (len(sids), 1))
for rep in range(4):
    t0 = time.perf_counter()
    m, v, st = ctx.predict_online(sids, thetas)
    print(os.environ.get("MEDGP_GRAPHS", "default"), "call", rep, round((time.perf_counter() - t0) * 1e3, 2), "ms", int((st != 0).sum()))
