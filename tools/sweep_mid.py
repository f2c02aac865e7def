"""NLML+gradient wall time per call for mid-size chunks, to place the dataflow-kernel policy.
usage: MEDGP_FLOW=-1|3 python tools/sweep_mid.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth
Q, D, R = 5, 24, 8
row = []
for n, batches in ((500, (96, 160, 256)), (900, (96, 192)), (1500, (2, 3, 5, 8, 16, 32)), (2500, (2, 3, 5, 8))):
    ctx = api.Context(Q, D, R, workspace_bytes=24 << 30)
    sid = ctx.add_series(*synth.make_patient(D, n, seed=1, T=240.0 * max(1.0, n / 500.0)))
    for batch in batches:
        thetas = synth.init_hyp_lmc_sm(Q, D, R, batch, seed=718)
        th = ctx.pinned(thetas.shape); th[...] = thetas
        outs = (ctx.pinned((batch,)), ctx.pinned((batch, ctx.P)), ctx.pinned((batch,), np.int32))
        for _ in range(3):
            ctx.nlml_grad([sid] * batch, th, True, out=outs)
        reps = 8
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.nlml_grad([sid] * batch, th, True, out=outs)
        row.append((n, batch, round((time.perf_counter() - t0) / reps * 1e3, 3)))
    ctx.close()
print(os.environ.get("MEDGP_FLOW", "auto"), row)
