"""Pins the FP64 oracle (oracle/medgp_oracle.c): against the committed outputs of the compiled
reference (tests/golden/golden.json, float arithmetic -> float tolerances), against central
finite differences, and through invariances the model has.  CPU only."""
import json
import os

import numpy as np
import pytest

from medgp_b200 import synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))

# the reference stores K, L, L^-1, W, alpha in float (c_inference_exact.cpp:66-68): its own
# noise floor vs FP64 is ~4e-7 on NLML and ~7e-6 on gradients
TOL_NLML_REF = 2e-6
TOL_GRAD_REF = 5e-5
TOL_PRED_REF = 2e-4


def case_inputs(c):
    meta, x, y = synth.make_patient(c["D"], c["n"], c["seed"])
    theta = synth.init_hyp_lmc_sm(c["Q"], c["D"], c["R"], 2, seed=c["theta_seed"])[1]
    return meta, x, y, theta


@pytest.mark.parametrize("idx", range(len(GOLD["eval"])))
def test_oracle_matches_reference_golden(oracle, idx):
    c = GOLD["eval"][idx]
    meta, x, y, theta = case_inputs(c)
    f, g, st = oracle.nlml_grad(c["Q"], c["D"], c["R"], meta, x, y, theta)
    assert st == 0
    assert abs(f - c["nlml"]) <= TOL_NLML_REF * abs(c["nlml"])
    gref = np.array(c["grad"])
    assert np.abs(g - gref).max() <= TOL_GRAD_REF * np.abs(gref).max()
    mu, var, _ = oracle.predict(c["Q"], c["D"], c["R"], meta, x, y, theta, np.array(c["star_meta"]),
                                np.array(c["star_x"], dtype=np.float32))
    assert np.abs(mu - np.array(c["pred_mean"])).max() <= TOL_PRED_REF
    assert np.abs(var - np.array(c["pred_var"])).max() <= TOL_PRED_REF


@pytest.mark.parametrize("Q,D,R,n", [(2, 3, 2, 60), (3, 4, 1, 90), (1, 1, 1, 25)])
def test_collapsed_gradient_equals_literal_and_finite_differences(oracle, Q, D, R, n):
    meta, x, y = synth.make_patient(D, n, seed=n)
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=5)[0]
    f0, g0, _ = oracle.nlml_grad(Q, D, R, meta, x, y, theta, grad_mode=0)
    _, g1, _ = oracle.nlml_grad(Q, D, R, meta, x, y, theta, grad_mode=1)   # reference's P dense dK passes
    assert np.abs(g0 - g1).max() <= 1e-12 * np.abs(g0).max()
    h = 1e-5
    for i in range(len(theta)):
        tp, tm = theta.copy(), theta.copy()
        tp[i] += h
        tm[i] -= h
        fd = (oracle.nlml_grad(Q, D, R, meta, x, y, tp, want_grad=False)[0]
              - oracle.nlml_grad(Q, D, R, meta, x, y, tm, want_grad=False)[0]) / (2 * h)
        assert abs(fd - g0[i]) <= 2e-6 * max(1.0, np.abs(g0).max()), (i, fd, g0[i])


def test_invariances(oracle):
    Q, D, R, n = 2, 3, 2, 70
    meta, x, y = synth.make_patient(D, n, seed=9)
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=3)[0]
    f0, g0, _ = oracle.nlml_grad(Q, D, R, meta, x, y, theta)
    perm = np.random.default_rng(0).permutation(n)       # point order is irrelevant
    f1, g1, _ = oracle.nlml_grad(Q, D, R, meta[perm], x[perm], y[perm], theta)
    assert abs(f1 - f0) <= 1e-11 * abs(f0) and np.abs(g1 - g0).max() <= 1e-9 * np.abs(g0).max()
    flipped = theta.copy()                                # B = A A^T: A -> -A changes nothing
    flipped[D:D + Q * D * R] *= -1
    f2, g2, _ = oracle.nlml_grad(Q, D, R, meta, x, y, flipped)
    assert abs(f2 - f0) <= 1e-12 * abs(f0)
    assert np.abs(g2[D:D + Q * D * R] + g0[D:D + Q * D * R]).max() <= 1e-9 * np.abs(g0).max()
    K = oracle.gram(Q, D, R, meta, x, theta)
    assert np.array_equal(K, K.T) and np.linalg.eigvalsh(K).min() > 0


def test_truncated_pi_matters_at_1e9(oracle):
    """the reference's PI = 3.14159265 must be used; true pi moves the
    gradient by more than the 1e-9 budget."""
    Q, D, R, n = 2, 2, 2, 80
    meta, x, y = synth.make_patient(D, n, seed=1)
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=719)[0]
    _, g_ref, _ = oracle.nlml_grad(Q, D, R, meta, x, y, theta, pi=3.14159265)
    _, g_true, _ = oracle.nlml_grad(Q, D, R, meta, x, y, theta, pi=np.pi)
    assert np.abs(g_ref - g_true).max() / np.abs(g_ref).max() > 1e-10


def test_jitter_semantics(oracle):
    """duplicated points + vanishing noise: K is singular until sigma^2 is added again;
    hopeless cases report status -1 (c_inference_exact.cpp:99-111)."""
    n = 40
    meta = np.zeros(n, dtype=np.int32)
    x = np.repeat(np.linspace(1, 10, n // 2), 2).astype(np.float32)
    y = np.random.default_rng(3).standard_normal(n).astype(np.float32)
    theta = np.array([np.log(1e-9), 1.0, np.log(1 / 24.0), np.log(1 / (2 * 3.14159265 * 48.0)), np.log(1e-12)])
    f, g, st = oracle.nlml_grad(1, 1, 1, meta, x, y, theta)
    assert st == -1 or st > 0


def test_prior_terms(oracle):
    lp, dlp = oracle.prior(1, 0.3, 0.0, 2.0)
    assert abs(lp - (-0.09 / 4.0 - np.log(2 * 3.14159265 * 2.0) / 2)) < 1e-15 and abs(dlp + 0.15) < 1e-15
    lp, dlp = oracle.prior(2, -0.3, 0.0, 0.5)
    assert abs(lp - (-0.6 - np.log(1.0))) < 1e-15 and dlp == 2.0
    assert oracle.prior(2, 0.0, 0.0, 0.5)[1] == 0.0


def test_numpy_large_n_oracle_agrees_with_c_oracle(oracle):
    from oracle import oracle_np
    Q, D, R, n = 3, 4, 2, 150
    meta, x, y = synth.make_patient(D, n, seed=21)
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=8)[0]
    perm = np.random.default_rng(1).permutation(n)
    f0, g0, _ = oracle.nlml_grad(Q, D, R, meta, x, y, theta)
    f1, g1 = oracle_np.nlml_grad_np(Q, D, R, meta[perm], x[perm], y[perm], theta)
    assert abs(f1 - f0) <= 1e-11 * abs(f0)
    assert np.abs(g1 - g0).max() <= 1e-10 * np.abs(g0).max()


def test_prefix_cholesky_leave_one_out_identity(oracle):
    """The identity behind medgp_cuda_predict_online, checked on the CPU with the oracle's own
    matrices: with points in time order and G = [a, b) the time-stamp group of j, the reference's
    per-observation refit (train = earlier points + the rest of G, main_one_test.cpp:286-366) equals
        var_j = 1 / (K_b^-1)_jj,  mean_j = y_j - (K_b^-1 y)_j / (K_b^-1)_jj,
    where only X = inv(L_GG) of the full factor is needed:
        (K_b^-1)_jj = sum_{i>=j} X_ij^2,  (K_b^-1 y)_j = sum_{i>=j} X_ij z_i,  z = L^-1 y."""
    Q, D, R, n = 2, 3, 2, 48
    meta, x, y = synth.make_patient(D, n, seed=91)
    x = np.round(x, 0).astype(np.float32)             # shared time stamps
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=8)[0]
    order = np.argsort(x, kind="stable")
    meta, x, y = meta[order], x[order], y[order]
    K = oracle.gram(Q, D, R, meta, x, theta, add_noise=True)
    L = np.linalg.cholesky(K)
    z = np.linalg.solve(L, y.astype(np.float64))
    starts = [0] + [i for i in range(1, n) if x[i] != x[i - 1]] + [n]
    checked = 0
    for a, b in zip(starts[:-1], starts[1:]):
        X = np.linalg.inv(L[a:b, a:b])
        for j in range(a, b):
            tr = [i for i in range(b) if i != j]
            if not tr:
                continue
            cc = float(np.sum(X[j - a:, j - a] ** 2))
            u = float(np.sum(X[j - a:, j - a] * z[j:b]))
            mu, var, st = oracle.predict(Q, D, R, meta[tr], x[tr], y[tr], theta, meta[j:j + 1], x[j:j + 1])
            assert st == 0
            assert abs(1.0 / cc - var[0]) <= 1e-10 * var[0]
            assert abs((float(y[j]) - u / cc) - mu[0]) <= 1e-9 * max(1.0, abs(mu[0]))
            checked += 1
    assert checked >= n - 1 and len(starts) - 1 < n    # at least one shared time stamp was exercised
