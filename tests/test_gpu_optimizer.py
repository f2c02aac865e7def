"""GPU: the device-resident lock-step optimiser (medgp_cuda_scg_*, medgp_b200/csrc/scg.cuh) --
the reference's SCG as a state machine in HBM -- against the committed trajectories of the
UNMODIFIED reference optimisers (tests/golden/golden.json: scg / varem, produced by
oracle/ref/ref_scg.cpp), and against the FP64 oracle for the prior terms it fuses."""
import json
import os
import subprocess

import numpy as np
import pytest

from medgp_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "medgp_b200", "host")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
PI = 3.14159265


def tagged(text):
    out = {}
    for line in text.split("\n"):
        p = line.split()
        if len(p) == 2 and p[0] in ("calls", "loss", "x", "v", "t"):
            out.setdefault(p[0], []).append(float(p[1]))
    return out


def run(args):
    return subprocess.run(args, check=True, capture_output=True, text=True, timeout=600).stdout


@pytest.mark.parametrize("idx", range(len(GOLD["scg"])))
def test_device_scg_follows_reference(idx):
    """the device state machine, driven through the session's taps on the analytic objective the
    reference's c_optimizer_scg was run on: same number of objective calls, same iterate
    (differences: summation order of the dot products)"""
    c = GOLD["scg"][idx]
    out = tagged(run([os.path.join(HOST, "host_check"), "dscg", str(c["max_iteration"])] + [repr(v) for v in c["x0"]]))
    # the longest case ends at a converged point, where a line-search branch hinges on the last
    # bits of f (tree vs sequential summation): allow one call of difference there
    assert abs(int(out["calls"][0]) - c["calls"]) <= (1 if c["calls"] > 40 else 0)
    assert abs(out["loss"][0] - c["loss"]) <= 1e-10 * abs(c["loss"])
    assert np.abs(np.array(out["x"]) - np.array(c["x"])).max() <= 1e-7


@pytest.mark.parametrize("idx", range(len(GOLD["varem"])))
def test_device_scg_under_variational_em_follows_reference(idx):
    """variational-EM rounds on the host (closed-form updates, pruning) around device-resident
    SCG runs, against the reference's c_optimizer_varEM"""
    c = GOLD["varem"][idx]
    out = tagged(run([os.path.join(HOST, "host_check"), "dvarem", str(c["max_iteration"]), str(c["sub_iter"]),
                      str(c["Q"]), str(c["D"]), str(c["R"]), "0.01", "0.01"] + [repr(v) for v in c["x0"]]))
    assert abs(int(out["calls"][0]) - c["calls"]) <= 2   # see tests/test_host_logic.py: branches on the last bits of f
    assert abs(out["loss"][0] - c["loss"]) <= 1e-9 * abs(c["loss"])
    assert np.abs(np.array(out["x"]) - np.array(c["x"])).max() <= 1e-5
    assert np.abs(np.array(out["v"]) - np.array(c["v"])).max() <= 1e-6 * max(1.0, np.abs(c["v"]).max())
    assert [int(v) for v in out["t"]] == c["t"]


def prior_terms(theta, ptype, pexp, ppar, D, nA):
    """numpy restatement of inference/c_inference_prior.cpp:59-150 + prior/c_prior.cpp:383-421"""
    lp_sum, dg, clamp = 0.0, np.zeros_like(theta), np.zeros(len(theta), dtype=bool)
    for k in range(len(theta)):
        t = ptype[k]
        if t < 0:
            continue
        if t == 0:
            clamp[k] = True
            continue
        h = theta[k] if D <= k < D + nA else np.exp(theta[k])
        p0, p1 = float(ppar[k, 0]), float(ppar[k, 1])
        if t == 1:
            lp, dlp = -(h - p0) ** 2 / (2 * p1) - np.log(2 * PI * p1) / 2, -(h - p0) / p1
        else:
            lp, dlp = -abs(h - p0) / p1 - np.log(2 * p1), (0.0 if h == p0 else -np.sign(h - p0) / p1)
        lp_sum += lp
        dg[k] = h * dlp if pexp[k] else dlp
    return lp_sum, dg, clamp


def test_device_scg_on_the_gp_objective_with_prior_terms(oracle):
    """First two super-steps on real series: the objective after the first evaluation is
    NLML - sum log p, and the second probe point is X - g / (1 + |g|^2) with the prior-adjusted
    (and clamped) gradient -- both against the oracle + a numpy restatement of the prior terms.
    Then a full run: every instance spends exactly its budget and does not get worse."""
    from medgp_b200 import api
    Q, D, R = 2, 3, 2
    nA = Q * D * R
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 30)
    P = ctx.P
    sizes = [70, 150, 333]
    series = [synth.make_patient(D, n, seed=40 + n) for n in sizes]
    sids = [ctx.add_series(*s) for s in series]
    theta0 = synth.init_hyp_lmc_sm(Q, D, R, len(sizes), seed=5)
    # hierarchical-gamma-like table: normal(0, psi) on A, laplace(0, 0.01) on kappa, one clamped A entry
    ptype = np.full((len(sizes), P), -1, dtype=np.int8)
    pexp = np.zeros((len(sizes), P), dtype=np.int8)
    ppar = np.zeros((len(sizes), P, 2), dtype=np.float32)
    ptype[:, D:D + nA] = 1
    ppar[:, D:D + nA, 1] = np.linspace(0.5, 1.5, nA, dtype=np.float32)
    ptype[:, D + nA + 2 * Q:] = 2
    pexp[:, D + nA + 2 * Q:] = 1
    ppar[:, D + nA + 2 * Q:, 1] = 0.01
    ptype[1, D + 3] = 0
    ses = ctx.scg_session(len(sizes))
    # budget -3 = two evaluations: the reference's counter advances twice per iteration
    # (c_optimizer_scg.cpp:73,88,114), a quirk the state machine keeps
    ses.start(sids, theta0, -3, ptype, pexp, ppar)
    assert ses.run(1) == len(sizes)
    pts, wants = ses.points()
    assert wants.all()
    for b, (meta, x, y) in enumerate(series):
        f0, g0, st0 = oracle.nlml_grad(Q, D, R, meta, x, y, theta0[b])
        lp, dg, clamp = prior_terms(theta0[b], ptype[b], pexp[b], ppar[b], D, nA)
        g = g0 - dg
        g[clamp] = 0.0
        expect = theta0[b] - g / (1.0 + g @ g)
        assert np.abs(pts[b] - expect).max() <= 1e-9 * max(1.0, np.abs(expect).max())
    assert ses.run(1) == 0          # two evaluations: budget spent
    best, loss, evals = ses.result()
    assert (evals == 2).all()
    for b, (meta, x, y) in enumerate(series):
        # the best point is theta0 or the second probe; its objective includes the prior terms
        cand = []
        for th in (theta0[b], pts[b]):
            f = oracle.nlml_grad(Q, D, R, meta, x, y, th, want_grad=False)[0]
            cand.append(f - prior_terms(th, ptype[b], pexp[b], ppar[b], D, nA)[0])
        assert abs(loss[b] - min(cand)) <= 1e-9 * abs(min(cand))
    # ---- a full run, mixed budgets, no priors: budgets are honoured exactly, instances finish
    #      at different super-steps, nobody ends above its starting objective
    budgets = np.array([-7, -25, -12], dtype=np.int32)
    ses.start(sids, theta0, budgets)
    left, rounds = len(sizes), 0
    while left:
        left = ses.run(5)
        rounds += 1
    best, loss, evals = ses.result()
    # never more evaluations than the budget, and the counter quirk eats at most half of it
    assert (evals <= -budgets).all() and (2 * evals >= -budgets).all()
    assert 5 * (rounds - 1) < evals.max() <= 5 * rounds
    for b, (meta, x, y) in enumerate(series):
        f_start = oracle.nlml_grad(Q, D, R, meta, x, y, theta0[b], want_grad=False)[0]
        f_best = oracle.nlml_grad(Q, D, R, meta, x, y, best[b], want_grad=False)[0]
        assert abs(f_best - loss[b]) <= 1e-9 * abs(f_best) and f_best <= f_start
    ses.close()
    ctx.close()
