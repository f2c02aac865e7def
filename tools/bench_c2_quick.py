"""C2 shape (patients x n points, one theta each): wall time per NLML+gradient call through the host ABI
with page-locked buffers, after warm-up.  usage: python tools/bench_c2_quick.py [patients] [n] [calls]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth
patients = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 30
Q, D, R = 5, 24, 8
ctx = api.Context(Q, D, R, workspace_bytes=16 << 30)
sids = [ctx.add_series(*synth.make_patient(D, n, seed=i)) for i in range(patients)]
thetas = synth.init_hyp_lmc_sm(Q, D, R, patients, seed=718)
th = ctx.pinned(thetas.shape); th[...] = thetas
outs = (ctx.pinned((patients,)), ctx.pinned((patients, ctx.P)), ctx.pinned((patients,), np.int32))
for _ in range(5):
    ctx.nlml_grad(sids, th, True, out=outs)
t0 = time.perf_counter()
for _ in range(calls):
    f, g, st = ctx.nlml_grad(sids, th, True, out=outs)
ms = (time.perf_counter() - t0) / calls * 1e3
print(f"{ms:.3f} ms/call, {patients / ms * 1e3:.0f} evals/s, failed {int((st < 0).sum())}")
