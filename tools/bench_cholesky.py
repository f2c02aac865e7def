"""FP64 % of peak of the blocked Cholesky on long series (BASELINE.json secondary metric).
Times NLML-only evaluations (prep + assemble + potrf + solve) of `count` initialisations of one
n-point, 24-feature, Q=5 series with per-stage CUDA events; potrf flops = n^3/3 per matrix.
usage: python tools/bench_cholesky.py [n] [count] [reps]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
want_grad = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
Q, D, R = 5, 24, 8
meta, x, y = synth.make_patient(D, n, seed=4000, T=1200.0)
thetas = synth.init_hyp_lmc_sm(Q, D, R, count, seed=4)
ctx = api.Context(Q, D, R, workspace_bytes=24 << 30)
sid = ctx.add_series(meta, x, y)
for _ in range(2):
    ctx.nlml_grad([sid] * count, thetas, want_grad)
import time
ctx.sync()
t0 = time.perf_counter()
for _ in range(reps):
    f, g, st = ctx.nlml_grad([sid] * count, thetas, want_grad)
wall = (time.perf_counter() - t0) / reps
ctx.stage_times(reset=True)
ctx.profile(True)
for _ in range(reps):
    ctx.nlml_grad([sid] * count, thetas, want_grad)
t = ctx.stage_times()
ms = {k: v["ms"] / reps for k, v in t.items() if k != "evals"}
potrf_tflops = count * n ** 3 / 3.0 / (ms["potrf"] * 1e-3) / 1e12
out = {"n": n, "count": count, "want_grad": want_grad, "wall_ms_multi_stream": wall * 1e3, "stage_ms_single_stream": ms,
       "potrf_tflops_single_stream": potrf_tflops, "evals_per_s": count / wall,
       "potrf_tflops_lower_bound_multi_stream": count * n ** 3 / 3.0 / wall / 1e12,
       "launches": {k: v["launches"] // reps for k, v in t.items() if k != "evals"}}
print(json.dumps(out))
assert (st == 0).all()
