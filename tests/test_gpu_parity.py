"""GPU parity tests: the CUDA path, called through the C ABI, against the FP64 oracle on the
same seeded inputs.  Tolerances are relative and stated per test (north_star: <= 1e-9)."""
import numpy as np
import pytest

from medgp_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-9          # north_star tolerance for NLML / gradients / predictions
RTOL_K = 1e-12       # covariance entries


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def api():
    from medgp_b200 import api as a
    return a


CASES = [
    # Q, D, R, n, seed
    (2, 2, 2, 80, 1),      # C1-like, one partial tile
    (2, 2, 1, 64, 2),      # exactly one tile
    (1, 1, 1, 37, 3),      # degenerate single output / single component
    (3, 4, 2, 130, 4),     # two tiles + 2 rows
    (5, 24, 8, 300, 5),    # C2 shape, small n
    (5, 24, 8, 500, 6),    # C2
    (8, 3, 2, 150, 7),     # Q = MEDGP_QMAX: widest kernel templates (gradient stage > 48 KB of dynamic smem)
    (7, 2, 1, 90, 8),
]


@pytest.mark.parametrize("Q,D,R,n,seed", CASES)
def test_matrices_nlml_grad(api, oracle, Q, D, R, n, seed):
    meta, x, y = synth.make_patient(D, n, seed)
    theta = synth.init_hyp_lmc_sm(Q, D, R, 2, seed=718 + seed)[1]
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 30)
    sid = ctx.add_series(meta, x, y)
    dbg = ctx.debug_matrices(sid, theta, n)
    K = oracle.gram(Q, D, R, meta, x, theta)
    assert rel(dbg["K"], K) <= RTOL_K
    alpha, L, logdet = oracle.fit(Q, D, R, meta, x, y, theta)
    assert rel(dbg["L"], L) <= 1e-10
    assert rel(dbg["alpha"], alpha) <= RTOL
    assert rel(dbg["Kinv"], np.linalg.inv(K)) <= RTOL
    f, g, st = ctx.nlml_grad([sid], theta[None, :], want_grad=True)
    f0, g0, st0 = oracle.nlml_grad(Q, D, R, meta, x, y, theta)
    assert st[0] == st0 == 0
    assert abs(f[0] - f0) <= RTOL * abs(f0)
    assert rel(g[0], g0) <= RTOL
    f2, _, _ = ctx.nlml_grad([sid], theta[None, :], want_grad=False)
    assert f2[0] == f[0]
    ctx.close()


@pytest.mark.parametrize("flow", [None, "3", "0"])
def test_ragged_batch_and_order_invariance(api, oracle, monkeypatch, flow):
    """sizes from 12 to 500 points in one call: 1 to 8 block rows in the same sub-chunk (the default
    schedule for such a chunk is right-looking with the dataflow inverse; flow = 3: dataflow
    factorisation too, whose ticket map must cope with the ragged block-row counts; 0: neither)"""
    if flow is not None:
        monkeypatch.setenv("MEDGP_FLOW", flow)
    Q, D, R = 3, 5, 2
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 30)
    rng = np.random.default_rng(0)
    sizes = [20, 64, 65, 200, 333, 129, 12, 500]
    thetas = synth.init_hyp_lmc_sm(Q, D, R, len(sizes), seed=11)
    sids, ref = [], []
    for k, n in enumerate(sizes):
        meta, x, y = synth.make_patient(D, n, 100 + k)
        perm = rng.permutation(n)          # arbitrary point order must not matter
        sids.append(ctx.add_series(meta[perm], x[perm], y[perm]))
        ref.append(oracle.nlml_grad(Q, D, R, meta, x, y, thetas[k]))
    f, g, st = ctx.nlml_grad(sids, thetas, want_grad=True)
    for k in range(len(sizes)):
        assert st[k] == 0
        assert abs(f[k] - ref[k][0]) <= RTOL * abs(ref[k][0])
        assert rel(g[k], ref[k][1]) <= RTOL
    # same series evaluated with several thetas in one call (random inits of one patient)
    f2, g2, _ = ctx.nlml_grad([sids[3]] * len(sizes), thetas, want_grad=True)
    meta, x, y = synth.make_patient(D, sizes[3], 103)
    for k in range(len(sizes)):
        f0, g0, _ = oracle.nlml_grad(Q, D, R, meta, x, y, thetas[k])
        assert abs(f2[k] - f0) <= RTOL * abs(f0)
        assert rel(g2[k], g0) <= RTOL
    ctx.close()


@pytest.mark.parametrize("rl", ["0", "1"])
def test_ragged_last_block_row(api, oracle, monkeypatch, rl):
    """The tile kernels do not issue the DMMAs of rows / columns beyond n in the last block row
    (8-row, 8-column, 4-deep granularity): every residue class that changes which sub-tiles are
    skipped, in both factorisation schedules, against the oracle -- NLML, gradient, and K^-1 itself."""
    monkeypatch.setenv("MEDGP_RL", rl)
    Q, D, R = 2, 3, 2
    sizes = [128 + r for r in (1, 4, 7, 8, 9, 31, 32, 33, 57, 63)] + [64 + 5, 3 * 64 + 40]
    series = [synth.make_patient(D, n, seed=300 + n) for n in sizes]
    thetas = synth.init_hyp_lmc_sm(Q, D, R, len(sizes), seed=21)
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 30)
    sids = [ctx.add_series(*s) for s in series]
    f, g, st = ctx.nlml_grad(sids, thetas, True)
    for k, n in enumerate(sizes):
        f0, g0, _ = oracle.nlml_grad(Q, D, R, *series[k], thetas[k])
        assert st[k] == 0 and abs(f[k] - f0) <= RTOL * abs(f0) and rel(g[k], g0) <= RTOL, n
    for k in (2, 5, 9):
        dbg = ctx.debug_matrices(sids[k], thetas[k], sizes[k])
        K = oracle.gram(Q, D, R, series[k][0], series[k][1], thetas[k])
        assert rel(dbg["Kinv"], np.linalg.inv(K)) <= RTOL
    ctx.close()


def test_duplicate_timestamps_and_two_point_features(api, oracle):
    Q, D, R = 2, 3, 2
    meta = np.array([0, 0, 1, 1, 1, 2, 2], dtype=np.int32)
    x = np.array([1.0, 5.0, 1.0, 5.0, 9.5, 5.0, 9.5], dtype=np.float32)   # r = 0 pairs across features
    y = np.array([0.3, -0.2, 1.0, 0.1, -0.7, 0.4, 0.0], dtype=np.float32)
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=5)[0]
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 28)
    sid = ctx.add_series(meta, x, y)
    f, g, st = ctx.nlml_grad([sid], theta[None], True)
    f0, g0, _ = oracle.nlml_grad(Q, D, R, meta, x, y, theta)
    assert abs(f[0] - f0) <= RTOL * abs(f0)
    assert rel(g[0], g0) <= RTOL
    ctx.close()


def test_jitter_path(api, oracle):
    """Tiny noise + duplicated points make K numerically singular: the reference adds sigma^2
    again and refactors (c_inference_exact.cpp:99-108); status counts the additions.  In this
    regime the matrix is at the edge of FP64 (condition ~1e16), so two correct implementations
    need not fail on the same attempt and the values are meaningless: only the contract is
    checked -- status in {-1, 0..10}, NaN exactly when -1, and a clear pass once the noise is
    comfortable.  Value parity of the jitter path is test_jitter_success_parity below."""
    Q, D, R = 1, 1, 1
    n = 40
    meta = np.zeros(n, dtype=np.int32)
    x = np.repeat(np.linspace(1, 10, n // 2), 2).astype(np.float32)
    y = np.random.default_rng(3).standard_normal(n).astype(np.float32)
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 28)
    sid = ctx.add_series(meta, x, y)
    seen = set()
    for log10_sigma in (-9.0, -8.0, -7.875, -7.75, -7.625, -7.5, -7.0, -5.0):
        theta = np.array([log10_sigma * np.log(10), 1.0, np.log(1 / 24.0), np.log(1 / (2 * 3.14159265 * 48.0)), np.log(1e-12)])
        f, g, st = ctx.nlml_grad([sid], theta[None], True)
        f0, g0, st0 = oracle.nlml_grad(Q, D, R, meta, x, y, theta)
        assert -1 <= st[0] <= 10
        assert np.isnan(f[0]) == (st[0] < 0) and np.isnan(g[0]).all() == (st[0] < 0)
        if log10_sigma <= -9.0:
            assert st[0] == st0 == -1      # 11 sigma^2 = 1e-17: hopeless in both
        if log10_sigma >= -5.0:
            assert st[0] == st0 == 0
            assert rel(f[0], f0) <= 1e-6   # condition ~1e10
        seen.add(int(st[0]))
    assert -1 in seen and 0 in seen
    ctx.close()


@pytest.mark.parametrize("attempts", [1, 3])
@pytest.mark.parametrize("retry", ["device", "host"])
def test_jitter_success_parity(api, oracle, monkeypatch, attempts, retry):
    """The jitter-success branch (0 < status <= 10, c_inference_exact.cpp:99-108) with values:
    the first `attempts` factorisations of every evaluation are declared failed in BOTH the CUDA
    library and the oracle, on well-conditioned series, so K + (1 + attempts) sigma^2 is what
    gets factored, the gradient uses W of the jittered K with the un-jittered dK, and NLML,
    gradient and predictions must agree to 1e-9 with status == attempts.  `device`: retries run
    inside the launch sequence (graph WHILE node), through the host ABI and through the
    device-resident ABI; `host`: the host-driven rounds (MEDGP_DEVICE_RETRY=0)."""
    Q, D, R = 2, 3, 2
    if retry == "host":
        monkeypatch.setenv("MEDGP_DEVICE_RETRY", "0")
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 30)
    ctx.force_fail(attempts)
    sizes = [150, 333, 64, 500]
    series = [synth.make_patient(D, n, seed=900 + n) for n in sizes]
    thetas = synth.init_hyp_lmc_sm(Q, D, R, len(sizes), seed=11)
    sids = [ctx.add_series(*s) for s in series]
    f, g, st = ctx.nlml_grad(sids, thetas, True)
    f_nograd, _, st_nograd = ctx.nlml_grad(sids, thetas, False)
    # device-resident entry point
    B, P = len(sids), ctx.P
    d_theta, d_nlml, d_grad, d_status = ctx.malloc(B * P * 8), ctx.malloc(B * 8), ctx.malloc(B * P * 8), ctx.malloc(B * 4)
    ctx.h2d(d_theta, thetas)
    for _ in range(2):  # second call replays the cached graph
        ctx.nlml_grad_device(np.array(sids, dtype=np.int32), d_theta, True, d_nlml, d_grad, d_status)
    ctx.sync()
    f_dev, g_dev, st_dev = np.empty(B), np.empty((B, P)), np.empty(B, dtype=np.int32)
    ctx.d2h(f_dev, d_nlml); ctx.d2h(g_dev, d_grad); ctx.d2h(st_dev, d_status)
    # predictions ride on the same factorisation
    star_m = np.array([0, 2], dtype=np.int32)
    star_x = np.array([10.5, 77.25], dtype=np.float32)
    mean, var, st_p = ctx.predict(sids[:2], thetas[:2], [0, 2, 4], np.tile(star_m, 2), np.tile(star_x, 2))
    oracle.force_fail(attempts)
    try:
        for b, (meta, x, y) in enumerate(series):
            f0, g0, st0 = oracle.nlml_grad(Q, D, R, meta, x, y, thetas[b])
            assert st0 == attempts
            assert st[b] == attempts and st_nograd[b] == attempts and st_dev[b] == attempts
            assert rel(f[b], f0) <= 1e-9 and rel(f_nograd[b], f0) <= 1e-9 and rel(f_dev[b], f0) <= 1e-9
            assert np.abs(g[b] - g0).max() <= 1e-9 * np.abs(g0).max()
            assert np.abs(g_dev[b] - g0).max() <= 1e-9 * np.abs(g0).max()
            if b < 2:
                m0, v0, stp0 = oracle.predict(Q, D, R, meta, x, y, thetas[b], star_m, star_x)
                assert st_p[b] == attempts == stp0
                assert np.abs(mean[2 * b:2 * b + 2] - m0).max() <= 1e-9 * max(1.0, np.abs(m0).max())
                assert np.abs(var[2 * b:2 * b + 2] - v0).max() <= 1e-9 * np.abs(v0).max()
        # the jittered evaluation differs from the plain one by far more than the tolerance
        oracle.force_fail(0)
        f_plain = oracle.nlml_grad(Q, D, R, *series[0], thetas[0], want_grad=False)[0]
        assert rel(f[0], f_plain) > 1e-4
    finally:
        oracle.force_fail(0)
    # more failed attempts than the reference allows: status -1, NaN
    ctx.force_fail(11)
    f, g, st = ctx.nlml_grad(sids[:2], thetas[:2], True)
    assert (st == -1).all() and np.isnan(f).all() and np.isnan(g).all()
    ctx.force_fail(0)
    f, g, st = ctx.nlml_grad(sids[:2], thetas[:2], True)
    assert (st == 0).all()
    for p in (d_theta, d_nlml, d_grad, d_status):
        ctx.free(p)
    ctx.close()


def test_device_path_needs_no_host_round_trip_and_mixed_statuses(api, oracle):
    """Consecutive device-resident calls are queued without host synchronisation; a batch in
    which some evaluations fail (and are retried on the device) leaves the others untouched."""
    Q, D, R = 1, 1, 1
    n = 40
    meta = np.zeros(n, dtype=np.int32)
    x_dup = np.repeat(np.linspace(1, 10, n // 2), 2).astype(np.float32)
    x_ok = np.linspace(1, 10, n).astype(np.float32)
    y = np.random.default_rng(3).standard_normal(n).astype(np.float32)
    hopeless = np.array([-9.0 * np.log(10), 1.0, np.log(1 / 24.0), np.log(1 / (2 * 3.14159265 * 48.0)), np.log(1e-12)])
    fine = np.array([np.log(0.3), 1.0, np.log(1 / 24.0), np.log(1 / (2 * 3.14159265 * 48.0)), np.log(0.1)])
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 28)
    s_dup, s_ok = ctx.add_series(meta, x_dup, y), ctx.add_series(meta, x_ok, y)
    sids = np.array([s_ok, s_dup, s_ok, s_dup], dtype=np.int32)
    thetas = np.stack([fine, hopeless, fine, fine])
    B, P = 4, ctx.P
    d_theta, d_nlml, d_grad, d_status = ctx.malloc(B * P * 8), ctx.malloc(B * 8), ctx.malloc(B * P * 8), ctx.malloc(B * 4)
    ctx.h2d(d_theta, thetas)
    for _ in range(6):  # more calls in flight than descriptor slots
        ctx.nlml_grad_device(sids, d_theta, True, d_nlml, d_grad, d_status)
    ctx.sync()
    f, st = np.empty(B), np.empty(B, dtype=np.int32)
    ctx.d2h(f, d_nlml); ctx.d2h(st, d_status)
    assert st[1] == -1 and np.isnan(f[1])
    for b in (0, 2, 3):
        f0, _, st0 = oracle.nlml_grad(Q, D, R, meta, x_ok if b != 3 else x_dup, y, thetas[b])
        assert st[b] == st0 == 0 and rel(f[b], f0) <= 1e-9
    for p in (d_theta, d_nlml, d_grad, d_status):
        ctx.free(p)
    ctx.close()


def test_export_factors_in_the_callers_order(api, oracle):
    """chol_alpha / chol_factor_inv, the out-parameters of the reference's compute_nlml
    (c_inference_exact.cpp:124-143), for a series in an arbitrary point order: alpha = K^-1 y and the
    lower-triangular L^-1 of THAT order, as floats (compared at float resolution with the oracle's
    factor), and NLML at 1e-9."""
    Q, D, R = 3, 4, 2
    rng = np.random.default_rng(12)
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 30)
    from medgp_b200.api import ORDER_GIVEN
    for n, seed in ((37, 1), (64, 2), (150, 3), (333, 4)):
        meta, x, y = synth.make_patient(D, n, seed=60 + seed)
        p = rng.permutation(n)
        meta, x, y = meta[p], x[p], y[p]          # not feature-major: the factor depends on the order
        theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=seed)[0]
        sid = ctx.add_series(meta, x, y, order=ORDER_GIVEN)
        alpha, linv, nlml, st = ctx.export_factors(sid, theta)
        a0, L0, _ = oracle.fit(Q, D, R, meta, x, y, theta)
        f0 = oracle.nlml_grad(Q, D, R, meta, x, y, theta, want_grad=False)[0]
        X0 = np.linalg.inv(L0)
        assert st == 0 and abs(nlml - f0) <= RTOL * abs(f0)
        assert np.abs(alpha - a0).max() <= 2e-7 * np.abs(a0).max()
        assert np.abs(linv - X0).max() <= 2e-7 * np.abs(X0).max()
        assert (np.triu(linv, 1) == 0).all()
        # NLML and predictions are available on such a series, gradients are not
        f, _, _ = ctx.nlml_grad([sid], theta[None], False)
        assert abs(f[0] - f0) <= RTOL * abs(f0)
        with pytest.raises(api.MedgpError):
            ctx.nlml_grad([sid], theta[None], True)
    # a FEATURE-ordered series has no exportable factor (its internal order is the library's)
    sid = ctx.add_series(meta, x, y)
    with pytest.raises(api.MedgpError):
        ctx.export_factors(sid, theta)
    ctx.close()


def test_changing_the_model_on_a_live_context(api, oracle):
    """medgp_cuda_model on a context that already captured launch sequences: the graphs bake in the
    model (dimensions by value, Q-templated kernels, shared-memory sizes), so none may survive.
    The same series evaluated under three models in a row -- including larger -> smaller Q with
    identically shaped batches, the case a stale graph would silently answer."""
    meta, x, y = synth.make_patient(2, 150, seed=31)
    ctx = api.Context(3, 2, 2, workspace_bytes=1 << 29)
    for (Q, D, R) in ((3, 2, 2), (1, 2, 2), (3, 2, 1), (3, 2, 2)):
        ctx.clear_series()
        ctx.set_model(Q, D, R)
        sids = [ctx.add_series(meta, x, y) for _ in range(3)]
        thetas = synth.init_hyp_lmc_sm(Q, D, R, 3, seed=8)
        for _ in range(3):  # first sighting: plain launches; second: captured; third: replayed
            f, g, st = ctx.nlml_grad(sids, thetas, True)
        for b in range(3):
            f0, g0, _ = oracle.nlml_grad(Q, D, R, meta, x, y, thetas[b])
            assert st[b] == 0 and abs(f[b] - f0) <= RTOL * abs(f0) and rel(g[b], g0) <= RTOL
    with pytest.raises(api.MedgpError):
        ctx.set_model(2, 2, 2)   # series still uploaded
    ctx.close()


def test_direct_captured_and_replayed_sequences_agree(api, monkeypatch):
    """A batch structure is issued as plain launches at its first sighting, captured into a CUDA graph
    at the second and replayed from the third call on (host-buffer entry points); MEDGP_LAZY_CAPTURE=0
    captures at once.  All of them must return the same bits, jitter rounds included."""
    Q, D, R = 3, 4, 2
    sizes = [70, 200, 333, 64, 129]
    series = [synth.make_patient(D, n, seed=700 + n) for n in sizes]
    thetas = synth.init_hyp_lmc_sm(Q, D, R, len(sizes), seed=3)
    results = []
    for lazy in ("1", "0"):
        monkeypatch.setenv("MEDGP_LAZY_CAPTURE", lazy)
        ctx = api.Context(Q, D, R, workspace_bytes=1 << 30)
        ctx.force_fail(1)  # every evaluation needs one jitter round: host-driven when direct, in-graph otherwise
        sids = [ctx.add_series(*s) for s in series]
        for _ in range(3):
            f, g, st = ctx.nlml_grad(sids, thetas, True)
            results.append((f.copy(), g.copy(), st.copy()))
        mean, var, _ = ctx.predict(sids[:2], thetas[:2], [0, 1, 2], series[0][0][:2], series[0][1][:2] + 0.5)
        results.append((mean.copy(), var.copy(), np.zeros(1)))
        ctx.close()
    for k in range(1, 3):
        for a, b in zip(results[0], results[k]):
            assert np.array_equal(a, b)
    for a, b in zip(results[0], results[4]):
        assert np.array_equal(a, b)
    assert np.array_equal(results[3][0], results[7][0]) and np.array_equal(results[3][1], results[7][1])
    assert (results[0][2] == 1).all()


def test_randomised_schedules_short_soak():
    """tools/fuzz_parity.py for a few seconds: random shapes, ragged batches, forced jitter rounds and
    random combinations of every scheduling switch, against the oracle at 1e-9 (own process: the
    switches are environment variables read when a context is created)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "8", "7"], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0 and "fuzz ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_argument_checks(api):
    Q, D, R = 2, 2, 1
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 28)
    meta, x, y = synth.make_patient(D, 30, seed=1)
    from medgp_b200.api import ORDER_TIME
    s_time = ctx.add_series(meta, x, y, order=ORDER_TIME)
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=1)
    f, _, st = ctx.nlml_grad([s_time], theta, False)          # NLML works on a time-ordered series
    assert st[0] == 0 and np.isfinite(f[0])
    with pytest.raises(api.MedgpError):                        # gradients do not: host entry point ...
        ctx.nlml_grad([s_time], theta, True)
    d = [ctx.malloc(ctx.P * 8), ctx.malloc(8), ctx.malloc(ctx.P * 8), ctx.malloc(4)]
    ctx.h2d(d[0], theta)
    with pytest.raises(api.MedgpError):                        # ... and device-resident entry point alike
        ctx.nlml_grad_device(np.array([s_time], dtype=np.int32), d[0], True, d[1], d[2], d[3])
    with pytest.raises(api.MedgpError):
        ctx.nlml_grad([12345], theta, False)                   # unknown series
    with pytest.raises(api.MedgpError):
        ctx.add_series(np.array([0, 5], dtype=np.int32), x[:2], y[:2])   # feature slot out of range
    f, g, st = ctx.nlml_grad([], np.zeros((0, ctx.P)), True)   # empty batch
    assert len(f) == 0
    for p in d:
        ctx.free(p)
    ctx.close()


def test_predict(api, oracle):
    Q, D, R = 3, 4, 2
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 30)
    thetas = synth.init_hyp_lmc_sm(Q, D, R, 3, seed=21)
    sids, offs, ms, xs, ref_m, ref_v = [], [0], [], [], [], []
    for k, (n, m) in enumerate([(50, 1), (200, 3), (131, 6)]):
        meta, x, y = synth.make_patient(D, n, 300 + k)
        rng = np.random.default_rng(k)
        mstar = rng.integers(0, D, m).astype(np.int32)
        xstar = rng.uniform(0.5, 260.0, m).astype(np.float32)
        xstar[0] = x[3]   # a test time that coincides with a training time
        sids.append(ctx.add_series(meta, x, y))
        offs.append(offs[-1] + m)
        ms.append(mstar)
        xs.append(xstar)
        mu, var, _ = oracle.predict(Q, D, R, meta, x, y, thetas[k], mstar, xstar)
        ref_m.append(mu)
        ref_v.append(var)
    mean, var, st = ctx.predict(sids, thetas, offs, np.concatenate(ms), np.concatenate(xs))
    assert (st == 0).all()
    assert rel(mean, np.concatenate(ref_m)) <= RTOL
    assert rel(var, np.concatenate(ref_v)) <= RTOL
    ctx.close()


def test_left_and_right_looking_factorisations_agree(api, oracle, monkeypatch):
    """n = 1100 (T = 18): one matrix takes the right-looking path, MEDGP_RL=0 forces the
    left-looking one; both must match the oracle."""
    Q, D, R, n = 3, 6, 2, 1100
    meta, x, y = synth.make_patient(D, n, seed=31)
    theta = synth.init_hyp_lmc_sm(Q, D, R, 1, seed=41)[0]
    from oracle import oracle_np
    f0, g0 = oracle_np.nlml_grad_np(Q, D, R, meta, x, y, theta)
    res = {}
    for rl in ("0", "1"):
        monkeypatch.setenv("MEDGP_RL", rl)
        ctx = api.Context(Q, D, R, workspace_bytes=2 << 30)
        sid = ctx.add_series(meta, x, y)
        f, g, st = ctx.nlml_grad([sid], theta[None], True)
        ctx.close()
        assert st[0] == 0
        assert abs(f[0] - f0) <= RTOL * abs(f0)
        assert rel(g[0], g0) <= RTOL
        res[rl] = (f[0], g[0])
    assert abs(res["0"][0] - res["1"][0]) <= 1e-12 * abs(f0)


def test_cohort_sweep_sizes(api):
    """C3-like ragged batch at full sizes (n up to 1500): checked through size-independent
    properties -- bitwise determinism across calls, invariance to point order and to the sign
    of A, and agreement with the numpy/LAPACK oracle on the largest series."""
    Q, D, R = 5, 24, 8
    rng = np.random.default_rng(7)
    sizes = [1500, 300, 777, 1210, 512, 1024]
    thetas = synth.init_hyp_lmc_sm(Q, D, R, len(sizes), seed=99)
    ctx = api.Context(Q, D, R, workspace_bytes=4 << 30)
    pats = [synth.make_patient(D, n, seed=500 + k) for k, n in enumerate(sizes)]
    sids = [ctx.add_series(*p) for p in pats]
    f1, g1, st = ctx.nlml_grad(sids, thetas, True)
    f2, g2, _ = ctx.nlml_grad(sids, thetas, True)
    assert (st == 0).all() and np.array_equal(f1, f2) and np.array_equal(g1, g2)
    shuffled = []
    for m, x, y in pats:
        p = rng.permutation(len(x))
        shuffled.append(ctx.add_series(m[p], x[p], y[p]))
    flipped = thetas.copy()
    flipped[:, D:D + Q * D * R] *= -1
    f3, g3, _ = ctx.nlml_grad(shuffled, flipped, True)
    assert np.abs(f3 - f1).max() <= 1e-10 * np.abs(f1).max()
    gA = slice(D, D + Q * D * R)
    for k in range(len(sizes)):
        assert rel(-g3[k, gA], g1[k, gA]) <= 1e-8 and rel(g3[k, :D], g1[k, :D]) <= 1e-8
    from oracle import oracle_np
    f0, g0 = oracle_np.nlml_grad_np(Q, D, R, *pats[0], thetas[0])
    assert abs(f1[0] - f0) <= RTOL * abs(f0) and rel(g1[0], g0) <= RTOL
    ctx.close()


def test_block_row_limit_of_the_dataflow_kernels():
    """n = 4096 (64 block rows: the dataflow kernels' limit) and n = 4097 / 4160 (beyond it: right-looking
    schedule) against the numpy oracle"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "edge_tmax.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "edge ok" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]


def test_long_stay_patient(api):
    """C4: n = 4000, 24 features, Q = 5, several initialisations of the same series in flight
    (right-looking blocked Cholesky path) against the numpy/LAPACK oracle."""
    Q, D, R, n = 5, 24, 8, 4000
    meta, x, y = synth.make_patient(D, n, seed=4000, T=1200.0)
    thetas = synth.init_hyp_lmc_sm(Q, D, R, 3, seed=4)
    ctx = api.Context(Q, D, R, workspace_bytes=8 << 30)
    sid = ctx.add_series(meta, x, y)
    f, g, st = ctx.nlml_grad([sid] * 3, thetas, True)
    assert (st == 0).all()
    from oracle import oracle_np
    f0, g0 = oracle_np.nlml_grad_np(Q, D, R, meta, x, y, thetas[1])
    assert abs(f[1] - f0) <= RTOL * abs(f0)
    assert rel(g[1], g0) <= RTOL
    fn, _, _ = ctx.nlml_grad([sid] * 3, thetas, False)
    assert np.array_equal(fn, f)
    ctx.close()


@pytest.mark.parametrize("env,batch", [
    ({"MEDGP_RL": "1"}, 6),                                 # right-looking potrf / trtri, look-ahead on a second stream
    ({"MEDGP_RL": "1", "MEDGP_LOOKAHEAD": "0"}, 6),         # right-looking, one stream
    ({"MEDGP_RL": "1", "MEDGP_RL_W": "1"}, 6),              # panels of 1, 2, 3 block columns (T = 6: ragged last panel)
    ({"MEDGP_RL": "1", "MEDGP_RL_W": "2"}, 6),
    ({"MEDGP_RL": "1", "MEDGP_RL_W": "3", "MEDGP_LOOKAHEAD": "0"}, 6),
    ({"MEDGP_RL": "1", "MEDGP_RL_W": "8"}, 6),              # one panel: left-looking inside, no trailing update
    ({"MEDGP_RL": "1", "MEDGP_FUSE_DIAG": "0"}, 6),         # right-looking with separate diagonal / panel kernels
    ({"MEDGP_RL": "1", "MEDGP_STREAMS": "1"}, 40),          # right-looking step kernel, 40 matrices on one stream
    ({"MEDGP_RL": "0"}, 6),                                 # left-looking, one launch per step (k_potrf_step)
    ({"MEDGP_FLOW": "3"}, 6),                               # dataflow kernels: one launch per factorisation / inverse, per-tile flags
    ({"MEDGP_FLOW": "3"}, 140),
    ({"MEDGP_FLOW": "3", "MEDGP_STREAMS": "1"}, 120),       # 120 matrices x 21 tile roles on ONE stream: several waves of waiting roles
    ({"MEDGP_FLOW": "1", "MEDGP_GRAPHS": "0"}, 40),         # dataflow factorisation only
    ({"MEDGP_FLOW": "2", "MEDGP_RL": "1"}, 6),              # right-looking factorisation + dataflow inverse (the default for few large matrices)
    ({"MEDGP_FLOW": "0", "MEDGP_RL": "1"}, 6),              # no dataflow kernels
    ({"MEDGP_RL": "0", "MEDGP_FUSE_DIAG": "0"}, 6),         # left-looking, separate kernels, folded diagonal update
    ({"MEDGP_RL": "0"}, 140),                               # left-looking, separate kernels, large batch
    ({"MEDGP_RL": "0", "MEDGP_STREAMS": "1", "MEDGP_GRAPHS": "0"}, 140),   # single stream, no CUDA graph
    ({"MEDGP_RL": "0", "MEDGP_CHAIN_DIAG": "1"}, 140),     # diagonal blocks factored inside the panel kernel
    ({"MEDGP_RL": "0", "MEDGP_STAGGER_US": "15", "MEDGP_GEMM_SMEM_PAD": "8192"}, 140),  # scheduling knobs
    # k_potrf_step with a grid of several waves (120 matrices x 6 block rows = 720 CTAs against
    # <= 444 resident, on ONE stream): roles are drawn from a ticket counter, so a panel CTA only
    # ever waits for a diagonal role that is already running, whatever order blocks are dispatched in
    ({"MEDGP_RL": "0", "MEDGP_STREAMS": "1"}, 120),
    ({"MEDGP_RL": "0", "MEDGP_DEVICE_RETRY": "0"}, 120),  # plain graph, no WHILE node
])
def test_every_factorisation_path(api, oracle, monkeypatch, env, batch):
    """all scheduling variants of kernel (2) give the oracle's numbers (n = 330: T = 6)"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    Q, D, R, n = 2, 4, 2, 330
    pats = [synth.make_patient(D, n - 7 * (k % 3), seed=900 + k) for k in range(batch)]
    thetas = synth.init_hyp_lmc_sm(Q, D, R, batch, seed=77)
    ctx = api.Context(Q, D, R, workspace_bytes=2 << 30)
    sids = [ctx.add_series(*p) for p in pats]
    f, g, st = ctx.nlml_grad(sids, thetas, True)
    f2, _, _ = ctx.nlml_grad(sids, thetas, False)
    ctx.close()
    assert (st == 0).all() and np.array_equal(f, f2)
    for k in list(range(0, batch, max(1, batch // 5)))[:6]:
        f0, g0, _ = oracle.nlml_grad(Q, D, R, *pats[k], thetas[k])
        assert abs(f[k] - f0) <= RTOL * abs(f0)
        assert rel(g[k], g0) <= RTOL


def test_pinned_host_buffers_take_the_direct_copy_path(api):
    """Page-locked caller buffers (medgp_cuda_host_alloc) are copied by DMA without staging;
    the results must be bit-identical to the pageable path, also when jitter rounds re-run."""
    Q, D, R = 2, 3, 2
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 30)
    sizes = [50, 130, 64, 7]
    thetas = synth.init_hyp_lmc_sm(Q, D, R, len(sizes), seed=5)
    thetas[1, :D] = -40.0  # noise ~ 0 on a series with duplicated timestamps: needs jitter or fails
    sids = []
    for k, n in enumerate(sizes):
        meta, x, y = synth.make_patient(D, n, 300 + k)
        if k == 1:
            x[1::2] = x[0::2][: len(x[1::2])]
            meta[:] = 0
        sids.append(ctx.add_series(meta, x, y))
    f0, g0, s0 = ctx.nlml_grad(sids, thetas, True)
    th = ctx.pinned(thetas.shape)
    th[...] = thetas
    outs = (ctx.pinned((len(sizes),)), ctx.pinned((len(sizes), ctx.P)), ctx.pinned((len(sizes),), np.int32))
    f1, g1, s1 = ctx.nlml_grad(sids, th, True, out=outs)
    assert f1 is outs[0] and g1 is outs[1]
    assert np.array_equal(s0, s1)
    assert np.array_equal(f0, f1, equal_nan=True)
    assert np.array_equal(g0, g1, equal_nan=True)
    ctx.close()


def sliding_window_reference(oracle, Q, D, R, meta, x, y, theta):
    """The reference's imputation loop without updates (main_one_test.cpp:269-444), one oracle
    fit per observation: train = earlier points + same-timestamp points other than the target."""
    n = len(x)
    mean, var = np.zeros(n), np.zeros(n)
    for tt in np.unique(x):
        past = np.flatnonzero(x < tt)
        curr = np.flatnonzero(x == tt)
        for j in curr:
            tr = np.concatenate([past, curr[curr != j]])
            if len(tr) == 0:
                mean[j] = np.nan  # reference: "no training observations" branch, handled on the host
                continue
            mu, v, _ = oracle.predict(Q, D, R, meta[tr], x[tr], y[tr], theta, meta[j:j + 1], x[j:j + 1])
            mean[j], var[j] = mu[0], v[0]
    return mean, var


def test_online_imputation_one_factorisation_per_series(api, oracle):
    """medgp_cuda_predict_online (prefix Cholesky + leave-one-out inside a timestamp group)
    against the per-observation refits of the reference's loop."""
    Q, D, R = 2, 4, 2
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 30)
    thetas = synth.init_hyp_lmc_sm(Q, D, R, 3, seed=33)
    cases = []
    for k, n in enumerate([40, 150, 97]):
        meta, x, y = synth.make_patient(D, n, 500 + k)
        rng = np.random.default_rng(k)
        # several features measured at the same instant, as a lab panel is: groups of 1..4
        x = np.round(x, 0 if k == 1 else 1).astype(np.float32)
        perm = rng.permutation(n)
        cases.append((meta[perm], x[perm], y[perm]))
    sids = [ctx.add_series(m, x, y, order=api.ORDER_TIME) for m, x, y in cases]
    mean, var, st = ctx.predict_online(sids, thetas)
    assert (st == 0).all()
    for k, (m, x, y) in enumerate(cases):
        rm, rv = sliding_window_reference(oracle, Q, D, R, m, x, y, thetas[k])
        has = ~np.isnan(rm)
        assert has.sum() >= len(x) - 1
        assert rel(mean[k][has], rm[has]) <= RTOL
        assert rel(var[k][has], rv[has]) <= RTOL
        # a first point without any training data: zero mean, prior variance (B_ff + sigma_f^2)
        for j in np.flatnonzero(~has):
            assert mean[k][j] == 0.0
            _, v0, _ = oracle.predict(Q, D, R, m[:1], x[:1] + 1e9, y[:1], thetas[k], m[j:j + 1], x[j:j + 1])
            assert abs(var[k][j] - v0[0]) <= 1e-6 * v0[0]
    # gradients are refused on a time-ordered series; NLML is not
    f, _, s0 = ctx.nlml_grad(sids[:1], thetas[:1], want_grad=False)
    f_ref = oracle.nlml_grad(Q, D, R, *cases[0], thetas[0], want_grad=False)[0]
    assert s0[0] == 0 and abs(f[0] - f_ref) <= RTOL * abs(f_ref)
    with pytest.raises(api.MedgpError):
        ctx.nlml_grad(sids[:1], thetas[:1], want_grad=True)
    ctx.close()


def test_online_imputation_rejects_oversized_timestamp_groups(api):
    """More than 32 observations on one time stamp cannot use the one-factorisation path: the
    time-ordered upload is refused (the front-end then refits per observation)."""
    Q, D, R = 1, 2, 1
    ctx = api.Context(Q, D, R, workspace_bytes=1 << 28)
    n = 40
    meta = (np.arange(n) % D).astype(np.int32)
    x = np.full(n, 3.5, dtype=np.float32)
    y = np.linspace(-1, 1, n).astype(np.float32)
    with pytest.raises(api.MedgpError):
        ctx.add_series(meta, x, y, order=api.ORDER_TIME)
    sid = ctx.add_series(meta, x, y)  # feature order is unaffected
    assert sid >= 0
    ctx.close()
