"""Where a block column of k_potrf_flow goes (one n = 4000 matrix): globaltimer stamps of the diagonal
role and of the first panel role of every column, from a library built with -DMEDGP_X_TRACE.
usage: MEDGP_LIB=medgp_b200/alt/libmedgp_x_trace.so python tools/flow_trace.py"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth  # noqa: E402

Q, D, R, n = 5, 24, 8, 4000
meta, x, y = synth.make_patient(D, n, seed=4000, T=1200.0)
ctx = api.Context(Q, D, R, workspace_bytes=12 << 30)
sid = ctx.add_series(meta, x, y)
theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=4)
for _ in range(3):
    ctx.nlml_grad([sid], theta, False)
buf = np.zeros((64, 16), dtype=np.uint64)
assert ctx.lib.medgp_cuda_debug_flow_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
t = buf.astype(np.float64) / 1e3  # us
names = ["diag: last operand flag seen -> products done", "diag: products done -> factor done + published",
         "panel(k+1,k): X flag seen after diag publish", "panel: flag seen -> X in smem", "panel: second product + store",
         "panel: rhs update", "panel: publish", "next diag: sees flag(k+1,k) after panel publish"]
rows = []
for k in range(2, 61):
    d, p, dn = t[k], t[k], t[k + 1]
    rows.append([d[2] - d[1], d[3] - d[2], p[10] - d[3], p[11] - p[10], p[12] - p[11], p[13] - p[12], p[14] - p[13], dn[1] - p[14],
                 dn[1] - d[1]])
rows = np.array(rows)
for i, nm in enumerate(names):
    print(f"{nm:55s} median {np.median(rows[:, i]):6.2f} us   mean {rows[:, i].mean():6.2f}")
print(f"{'column period (flag(k,k-1) seen -> flag(k+1,k) seen)':55s} median {np.median(rows[:, 8]):6.2f} us   mean {rows[:, 8].mean():6.2f}")
