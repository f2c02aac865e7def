"""The dataflow kernels serve up to 64 block rows (n <= 4096): n = 4096 takes them, n = 4097 / 4160 fall back to the
right-looking schedule; both against the numpy oracle at 1e-9.  usage: python tools/edge_tmax.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from medgp_b200 import api, synth
from oracle import oracle_np
Q, D, R = 2, 4, 2
for n in (4096, 4097, 4160):
    meta, x, y = synth.make_patient(D, n, seed=n, T=1200.0)
    theta = synth.init_hyp_lmc_sm(Q, D, R, 2, seed=4)
    ctx = api.Context(Q, D, R, workspace_bytes=8 << 30)
    sid = ctx.add_series(meta, x, y)
    f, g, st = ctx.nlml_grad([sid, sid], theta, True)
    f0, g0 = oracle_np.nlml_grad_np(Q, D, R, meta, x, y, theta[1])
    rel = np.abs(g[1] - g0).max() / np.abs(g0).max()
    print(n, st, abs(f[1] - f0) / abs(f0), rel)
    assert st[1] == 0 and abs(f[1] - f0) <= 1e-9 * abs(f0) and rel <= 1e-9
    ctx.close()
print("edge ok")
