"""`count` evaluations of one n-point series (NLML only), a few calls: the workload for ncu captures of
the few-large-matrices path.  usage: python tools/longstay_one.py [count] [n] [calls] [want_grad]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth  # noqa: E402

count = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 2
grad = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
Q, D, R = 5, 24, 8
meta, x, y = synth.make_patient(D, n, seed=4000, T=1200.0)
ctx = api.Context(Q, D, R, workspace_bytes=12 << 30)
sid = ctx.add_series(meta, x, y)
thetas = synth.init_hyp_lmc_sm(Q, D, R, count, seed=4)
for _ in range(calls):
    f, g, st = ctx.nlml_grad([sid] * count, thetas, grad)
print(f, st)
