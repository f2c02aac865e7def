"""Randomised parity soak: random model shapes, ragged batches and scheduling switches, NLML + gradient
(+ a prediction) against the FP64 oracle at 1e-9.  Each round runs in a fresh context with a random
combination of the scheduling environment switches.
usage: python tools/fuzz_parity.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth  # noqa: E402
from oracle import oracle  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
SWITCHES = {"MEDGP_RL": ["0", "1", None], "MEDGP_FLOW": ["0", "1", "2", "3", None], "MEDGP_RL_W": ["1", "2", "3", "4", None],
            "MEDGP_LOOKAHEAD": ["0", None], "MEDGP_FUSE_DIAG": ["0", None], "MEDGP_STREAMS": ["1", "3", None],
            "MEDGP_GRAPHS": ["0", None], "MEDGP_LAZY_CAPTURE": ["0", None], "MEDGP_CHAIN_DIAG": ["1", None],
            "MEDGP_DEAL": ["0", None], "MEDGP_DEVICE_RETRY": ["0", None]}


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


t_end = time.time() + budget
rounds = evals = 0
worst = 0.0
while time.time() < t_end:
    env = {}
    for k, vals in SWITCHES.items():
        v = vals[rng.integers(len(vals))]
        os.environ.pop(k, None)
        if v is not None:
            os.environ[k] = env[k] = v
    Q, D, R = int(rng.integers(1, 6)), int(rng.integers(1, 7)), int(rng.integers(1, 4))
    batch = int(rng.choice([1, 2, 5, 9, 40, 150]))
    nmax = int(rng.choice([70, 200, 450, 700])) if batch <= 40 else 200
    sizes = rng.integers(max(2 * D, 3), nmax + 1, batch)
    npat = min(batch, 6)
    pats = [synth.make_patient(D, int(n), seed=int(rng.integers(1 << 30))) for n in sizes[:npat]]
    which = rng.integers(npat, size=batch)
    thetas = synth.init_hyp_lmc_sm(Q, D, R, batch, seed=int(rng.integers(1 << 30)))
    ctx = api.Context(Q, D, R, workspace_bytes=2 << 30)
    force = int(rng.choice([0, 0, 0, 1, 2]))
    if force:
        ctx.force_fail(force)
        oracle.force_fail(force)
    sids = [ctx.add_series(*p) for p in pats]
    for rep in range(int(rng.integers(1, 4))):  # repeated calls: direct -> captured -> replayed
        f, g, st = ctx.nlml_grad([sids[w] for w in which], thetas, True)
    check = rng.choice(batch, size=min(batch, 4), replace=False)
    for b in check:
        f0, g0, st0 = oracle.nlml_grad(Q, D, R, *pats[which[b]], thetas[b])
        ok = st[b] == st0 == force and abs(f[b] - f0) <= 1e-9 * abs(f0) and rel(g[b], g0) <= 1e-9
        worst = max(worst, abs(f[b] - f0) / abs(f0), rel(g[b], g0))
        if not ok:
            print("MISMATCH", dict(env=env, Q=Q, D=D, R=R, n=int(len(pats[which[b]][1])), batch=batch, sizes=[int(len(p[1])) for p in pats],
                                   status=(int(st[b]), int(st0)), f=(float(f[b]), float(f0)), grel=rel(g[b], g0)))
            sys.exit(1)
        evals += 1
    # predictions (cross-covariance columns ride along the factorisation as extra right-hand sides)
    npred = min(npat, 3)
    nstar = rng.integers(1, 4, npred)
    offs = np.concatenate([[0], np.cumsum(nstar)]).astype(np.int32)
    mstar = rng.integers(0, D, int(offs[-1])).astype(np.int32)
    xstar = rng.uniform(0.0, 250.0, int(offs[-1])).astype(np.float32)
    mean, var, stp = ctx.predict(sids[:npred], thetas[:npred], offs, mstar, xstar)
    for b in range(npred):
        sl = slice(offs[b], offs[b + 1])
        m0, v0, sp0 = oracle.predict(Q, D, R, *pats[b], thetas[b], mstar[sl], xstar[sl])
        ok = stp[b] == force and np.abs(mean[sl] - m0).max() <= 1e-9 * max(1.0, np.abs(m0).max()) and rel(var[sl], v0) <= 1e-9
        if not ok:
            print("PREDICTION MISMATCH", dict(env=env, Q=Q, D=D, R=R, n=int(len(pats[b][1])), status=int(stp[b]), mean=(mean[sl], m0), var=(var[sl], v0)))
            sys.exit(1)
        evals += 1
    if force:
        oracle.force_fail(0)
    ctx.close()
    rounds += 1
print(f"fuzz ok: {rounds} rounds, {evals} evaluations checked against the oracle, worst relative deviation {worst:.2e}")
