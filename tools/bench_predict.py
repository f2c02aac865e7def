"""Online one-step-ahead imputation throughput (BASELINE.json config 5, mean_wo_update mode of
main_one_test): every observation of a patient is predicted from all earlier observations plus
the other observations sharing its time stamp.
  refit  : one training set per observation, all of them batched through medgp_cuda_predict
           (what the reference does, one factorisation each)
  online : one factorisation per patient (medgp_cuda_predict_online)
usage: python tools/bench_predict.py [patients] [n_points]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth  # noqa: E402

patients = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
Q, D, R = 5, 24, 8
ctx = api.Context(Q, D, R, workspace_bytes=32 << 30)
theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=718)[0]
data = []
for p in range(patients):
    meta, x, y = synth.make_patient(D, n, seed=p)
    data.append((meta, np.round(x, 1).astype(np.float32), y))  # 0.1 h resolution: some shared time stamps

# ---- refit: one series per observation
t0 = time.perf_counter()
sids, offs, ms, xs, where = [], [0], [], [], []
for p, (meta, x, y) in enumerate(data):
    for j in range(n):
        tr = np.flatnonzero((x < x[j]) | ((x == x[j]) & (np.arange(n) != j)))
        if len(tr) == 0:
            continue
        sids.append(ctx.add_series(meta[tr], x[tr], y[tr]))
        ms.append(meta[j]); xs.append(x[j]); where.append((p, j))
        offs.append(len(ms))
t_up = time.perf_counter() - t0
thetas = np.tile(theta, (len(sids), 1))
ctx.predict(sids[:8], thetas[:8], offs[:9], np.array(ms[:8]), np.array(xs[:8], dtype=np.float32))  # warm-up
t0 = time.perf_counter()
mean, var, st = ctx.predict(sids, thetas, offs, np.array(ms), np.array(xs, dtype=np.float32))
t_refit = time.perf_counter() - t0
for s in sids:
    ctx.free_series(s)

# ---- online: one time-ordered series per patient
t0 = time.perf_counter()
osids = [ctx.add_series(m, x, y, order=api.ORDER_TIME) for m, x, y in data]
t_up2 = time.perf_counter() - t0
othetas = np.tile(theta, (patients, 1))
ctx.predict_online(osids[:1], othetas[:1])  # warm-up
t0 = time.perf_counter()
omean, ovar, ost = ctx.predict_online(osids, othetas)
t_online = time.perf_counter() - t0

dm = max(abs(omean[p][j] - mean[k]) / max(abs(mean[k]), 1e-300) for k, (p, j) in enumerate(where))
dv = max(abs(ovar[p][j] - var[k]) / var[k] for k, (p, j) in enumerate(where))
truth = np.array([data[p][2][j] for p, j in where])
cover = np.mean(np.abs(mean - truth) <= 1.96 * np.sqrt(var))
print(json.dumps({
    "patients": patients, "n_points": n, "predictions": len(sids),
    "refit": {"upload_s": t_up, "predict_s": t_refit, "predictions_per_s": len(sids) / t_refit,
              "failed": int((st < 0).sum())},
    "online": {"upload_s": t_up2, "predict_s": t_online, "predictions_per_s": len(sids) / t_online,
               "failed": int((ost < 0).sum())},
    "speedup": t_refit / t_online, "max_rel_diff_mean": float(dm), "max_rel_diff_var": float(dv),
    "ci95_coverage": float(cover)}))
