"""Run a batch several times and compare bitwise. usage: determinism_check.py [B] [n] [want_grad]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth
from oracle import oracle
Q, D, R = 5, 24, 8
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
wg = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
ctx = api.Context(Q, D, R, workspace_bytes=8 << 30)
pats = [synth.make_patient(D, n, seed=i) for i in range(B)]
sids = [ctx.add_series(*p) for p in pats]
thetas = synth.init_hyp_lmc_sm(Q, D, R, B, seed=718)
res = []
for rep in range(4):
    f, g, st = ctx.nlml_grad(sids, thetas, wg)
    res.append((f.copy(), None if g is None else g.copy()))
ref = [oracle.nlml_grad(Q, D, R, *pats[i], thetas[i], want_grad=False)[0] for i in range(min(B, 16))]
print("B", B, "n", n, "grad", wg, "streams", os.environ.get("MEDGP_STREAMS"))
for a in range(4):
    err = max(abs(res[a][0][i] - ref[i]) / abs(ref[i]) for i in range(len(ref)))
    print(" run", a, "max rel err vs oracle (first 16): %.3e" % err)
for a in range(3):
    df = np.abs(res[a][0] - res[3][0]) / np.abs(res[3][0])
    print(" run", a, "vs 3: max rel nlml diff %.3e, n_diff=%d" % (df.max(), (df > 0).sum()))
