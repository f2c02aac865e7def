"""Latency of ONE NLML+gradient evaluation through the host ABI (the shape of main_one_train's
SCG loop: batch = 1) and of small batches.  usage: python tools/bench_latency.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth  # noqa: E402

out = []
for (Q, D, R, n) in [(2, 2, 2, 100), (2, 2, 2, 300), (5, 24, 8, 500), (5, 24, 8, 1500), (5, 24, 8, 4000)]:
    ctx = api.Context(Q, D, R, workspace_bytes=8 << 30)
    sid = ctx.add_series(*synth.make_patient(D, n, seed=1, T=240.0 * max(1.0, n / 500.0)))
    for batch in (1, 5):
        if n == 4000 and batch > 1 and False:
            continue
        thetas = synth.init_hyp_lmc_sm(Q, D, R, batch, seed=718)
        th = ctx.pinned(thetas.shape)
        th[...] = thetas
        outs = (ctx.pinned((batch,)), ctx.pinned((batch, ctx.P)), ctx.pinned((batch,), np.int32))
        for _ in range(3):
            ctx.nlml_grad([sid] * batch, th, True, out=outs)
        reps = 20 if n < 4000 else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            f, g, st = ctx.nlml_grad([sid] * batch, th, True, out=outs)
        dt = (time.perf_counter() - t0) / reps
        out.append({"Q": Q, "D": D, "R": R, "n": n, "batch": batch, "ms_per_call": dt * 1e3, "status": int(st[0])})
    ctx.close()
print(json.dumps(out))
