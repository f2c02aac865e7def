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
names = ["panel(k+1,k): flag of block k seen -> X_kk in smem", "panel: second product + tile stored", "panel: rhs update",
         "panel: publish tile", "chained: last term from smem (+ wait for D')", "chained: factor block k+1, publish",
         "panel(k+2,k+1): sees the flag"]
rows = []
for k in range(2, 60):
    p, d, pn = t[k], t[k + 1], t[k + 1]
    rows.append([p[11] - p[10], p[12] - p[11], p[13] - p[12], p[14] - p[13], d[2] - p[14], d[3] - d[2], pn[10] - d[3], pn[10] - p[10]])
rows = np.array(rows)
for i, nm in enumerate(names):
    print(f"{nm:55s} median {np.median(rows[:, i]):6.2f} us   mean {rows[:, i].mean():6.2f}")
print(f"{'column period (flag of block k seen -> flag of block k+1 seen)':55s} median {np.median(rows[:, 7]):6.2f} us   mean {rows[:, 7].mean():6.2f}")
