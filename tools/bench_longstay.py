"""C4 (long-stay patient, n = 4000): NLML only and NLML+gradient with 1, 5 and 32 initialisations in
flight; wall time per call through the host ABI and the factorisation's share from per-stage events.
usage: python tools/bench_longstay.py [n]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
Q, D, R = 5, 24, 8
meta, x, y = synth.make_patient(D, n, seed=4000, T=1200.0)
ctx = api.Context(Q, D, R, workspace_bytes=12 << 30)
sid = ctx.add_series(meta, x, y)
out = {"n": n, "lookahead": os.environ.get("MEDGP_LOOKAHEAD", "1")}
for count in (1, 5, 32):
    thetas = synth.init_hyp_lmc_sm(Q, D, R, count, seed=4)
    for grad in (False, True):
        for _ in range(3):
            ctx.nlml_grad([sid] * count, thetas, grad)
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            f, g, st = ctx.nlml_grad([sid] * count, thetas, grad)
        ms = (time.perf_counter() - t0) / reps * 1e3
        ctx.stage_times(reset=True)
        ctx.profile(True)
        for _ in range(2):
            ctx.nlml_grad([sid] * count, thetas, grad)
        t = ctx.stage_times(reset=True)
        ctx.profile(False)
        potrf_ms = (t["potrf"]["ms"] + t["diag"]["ms"]) / 2
        flop = count * n ** 3 * (1.0 if grad else 1.0 / 3.0)
        out[f"{count}_{'grad' if grad else 'nlml'}"] = {
            "ms_per_call": ms, "tflops_call": flop / (ms * 1e-3) / 1e12, "potrf_ms_profiled": potrf_ms,
            "potrf_tflops_profiled": count * n ** 3 / 3.0 / (potrf_ms * 1e-3) / 1e12, "ok": bool((st == 0).all())}
print(json.dumps(out))
