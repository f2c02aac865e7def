"""Relative NLML / gradient error against the FP64 oracle on an ill-conditioned single-tile series
(duplicated time stamps, small noise: condition up to ~1e12) and on a well-conditioned multi-tile one;
run with MEDGP_LIB=<alternative build> to compare builds.  usage: python tools/accuracy_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medgp_b200 import api, synth  # noqa: E402
from oracle import oracle  # noqa: E402


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


Q, D, R = 1, 1, 1
for n in (40, 64, 200):
    meta = np.zeros(n, dtype=np.int32)
    x = np.repeat(np.linspace(1, 10, n // 2), 2).astype(np.float32)
    y = np.random.default_rng(3).standard_normal(n).astype(np.float32)
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 28)
    sid = ctx.add_series(meta, x, y)
    for log10_sigma in (-6.0, -5.0, -4.0, -3.0, -2.0, -1.0):
        theta = np.array([log10_sigma * np.log(10), 1.0, np.log(1 / 24.0), np.log(1 / (2 * 3.14159265 * 48.0)), np.log(1e-12)])
        f, g, st = ctx.nlml_grad([sid], theta[None], True)
        f0, g0, st0 = oracle.nlml_grad(Q, D, R, meta, x, y, theta)
        print(f"n={n:4d} sigma=1e{log10_sigma:+.0f} status {st[0]}/{st0} rel nlml {rel(f[0], f0):.2e} rel grad {rel(g[0], g0):.2e}")
    ctx.close()
Q, D, R = 5, 24, 8
for n in (500, 1500):
    meta, x, y = synth.make_patient(D, n, seed=11)
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=718)[0]
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 30)
    sid = ctx.add_series(meta, x, y)
    f, g, st = ctx.nlml_grad([sid], theta[None], True)
    f0, g0, st0 = oracle.nlml_grad(Q, D, R, meta, x, y, theta)
    print(f"C2-like n={n}: rel nlml {rel(f[0], f0):.2e} rel grad {rel(g[0], g0):.2e}")
    ctx.close()
