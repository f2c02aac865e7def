"""numpy/LAPACK restatement of the NLML + gradient for LARGE series (n in the thousands), where
the scalar C oracle would take minutes.  TEST INFRASTRUCTURE (same rules as medgp_oracle.c).
The Gram matrix comes from the C oracle (FP64, same formulas); factorisation and inverse use
LAPACK through scipy; the gradient is the collapsed block-sum form,
cross-checked against the C oracle on small cases in tests/test_oracle_pin.py."""
from __future__ import annotations

import numpy as np
import scipy.linalg as sl

from . import oracle

PI_REF = oracle.PI_REF


def nlml_grad_np(Q, D, R, meta, x, y, theta, want_grad=True, pi=PI_REF):
    meta = np.asarray(meta)
    order = np.argsort(meta, kind="stable")          # feature-major (results are order free)
    meta, x, y = meta[order], np.asarray(x, dtype=np.float32)[order], np.asarray(y, dtype=np.float32)[order]
    n = len(x)
    K = oracle.gram(Q, D, R, meta, x, theta, add_noise=True, pi=pi)
    cf = sl.cho_factor(K, lower=True, check_finite=False)
    yy = y.astype(np.float64)
    alpha = sl.cho_solve(cf, yy, check_finite=False)
    nlml = 0.5 * yy @ alpha + np.log(np.diag(cf[0])).sum() + n * np.log(2.0 * pi) / 2.0
    if not want_grad:
        return nlml, None
    W = sl.cho_solve(cf, np.eye(n), check_finite=False) - np.outer(alpha, alpha)
    cov = theta[D:]
    A = cov[:Q * D * R].reshape(Q, D, R)
    mu, v = np.exp(cov[Q * D * R:Q * D * R + Q]), np.exp(cov[Q * (D * R + 1):Q * (D * R + 2)])
    kappa = np.exp(cov[Q * (D * R + 2):]).reshape(Q, D)
    sigma = np.exp(theta[:D])
    counts = np.bincount(meta, minlength=D)
    present = np.nonzero(counts)[0]
    offs = np.concatenate([[0], np.cumsum(counts[present])[:-1]])

    def block_sums(Mx):
        S = np.zeros((D, D))
        S[np.ix_(present, present)] = np.add.reduceat(np.add.reduceat(Mx, offs, axis=0), offs, axis=1)
        return S

    t = x.astype(np.float64)
    tau = np.abs(t[:, None] - t[None, :])
    g = np.zeros(D + Q * (D * R + 2 + D))
    dW = np.zeros(D)
    np.add.at(dW, meta, np.diag(W))
    g[:D] = sigma ** 2 * dW
    for q in range(Q):
        B = A[q] @ A[q].T + np.diag(kappa[q])
        phi = 2.0 * pi * tau * mu[q]
        ex = np.exp(-2.0 * (pi * v[q]) ** 2 * tau ** 2)
        k = np.cos(phi) * ex
        Sk = block_sums(W * k)
        Sm = block_sums(W * (-phi * np.sin(phi) * ex))
        Sv = block_sums(W * (-4.0 * (pi * v[q]) ** 2 * tau ** 2 * k))
        g[D + q * D * R:D + (q + 1) * D * R] = (Sk @ A[q]).ravel()
        g[D + Q * D * R + q] = 0.5 * (B * Sm).sum()
        g[D + Q * (D * R + 1) + q] = 0.5 * (B * Sv).sum()
        g[D + Q * (D * R + 2) + q * D:D + Q * (D * R + 2) + (q + 1) * D] = 0.5 * kappa[q] * np.diag(Sk)
    return nlml, g
