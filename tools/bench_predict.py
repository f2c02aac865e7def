"""Online one-step-ahead prediction throughput (BASELINE.json config 5, mean_wo_update mode):
every observation of a patient is predicted from all earlier observations, all training
prefixes of all patients batched through medgp_cuda_predict.
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
t0 = time.perf_counter()
sids, offs, ms, xs, truth = [], [0], [], [], []
for p in range(patients):
    meta, x, y = synth.make_patient(D, n, seed=p)
    order = np.argsort(x, kind="stable")
    for k in range(3, n):                       # predict point order[k] from the k earlier ones
        past = order[:k]
        past = past[x[past] < x[order[k]]]
        if len(past) < 3:
            continue
        sids.append(ctx.add_series(meta[past], x[past], y[past]))
        ms.append(meta[order[k]]); xs.append(x[order[k]]); truth.append(y[order[k]])
        offs.append(len(ms))
t_up = time.perf_counter() - t0
thetas = np.tile(theta, (len(sids), 1))
t0 = time.perf_counter()
mean, var, st = ctx.predict(sids, thetas, offs, np.array(ms), np.array(xs, dtype=np.float32))
dt = time.perf_counter() - t0
cover = np.mean(np.abs(mean - np.array(truth)) <= 1.96 * np.sqrt(var))
print(json.dumps({"patients": patients, "n_points": n, "predictions": len(sids), "upload_s": t_up, "predict_s": dt,
                  "predictions_per_s": len(sids) / dt, "failed": int((st < 0).sum()), "ci95_coverage": float(cover)}))
