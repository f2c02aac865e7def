"""ctypes front-end of the TEST oracle (oracle/medgp_oracle.c) and of the compiled
reference driver oracle/_ref/ref_eval.  Test infrastructure: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmedgp_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_EVAL = os.path.join(REF_DIR, "ref_eval")
PI_REF = 3.14159265

_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "libmedgp_oracle.so"], check=True,
                   stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _prep(meta, x, y=None):
    meta = np.ascontiguousarray(meta, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    if y is not None:
        y = np.ascontiguousarray(y, dtype=np.float32)
    return meta, x, y


def force_fail(attempts):
    """Tests of the jitter path: the first `attempts` factorisations of every later evaluation
    count as failed (0 = off)."""
    lib().medgp_oracle_force_fail(int(attempts))


def nlml_grad(Q, D, R, meta, x, y, theta, want_grad=True, grad_mode=0, pi=PI_REF):
    """Returns (nlml, grad or None, status)."""
    meta, x, y = _prep(meta, x, y)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    P = D + Q * (D * R + 2 + D)
    assert theta.shape == (P,)
    nlml = ctypes.c_double(0.0)
    status = ctypes.c_int(0)
    grad = np.zeros(P, dtype=np.float64)
    rc = lib().medgp_oracle_nlml_grad(
        Q, D, R, ctypes.c_double(pi), len(x), _p(meta, ctypes.c_int32), _p(x, ctypes.c_float),
        _p(y, ctypes.c_float), _p(theta, ctypes.c_double), int(want_grad), int(grad_mode),
        ctypes.byref(nlml), _p(grad, ctypes.c_double), ctypes.byref(status))
    if rc != 0:
        return float("nan"), None, status.value
    return nlml.value, (grad if want_grad else None), status.value


def gram(Q, D, R, meta, x, theta, add_noise=True, pi=PI_REF):
    meta, x, _ = _prep(meta, x)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n = len(x)
    K = np.zeros((n, n), dtype=np.float64)
    lib().medgp_oracle_gram(Q, D, R, ctypes.c_double(pi), n, _p(meta, ctypes.c_int32),
                            _p(x, ctypes.c_float), _p(theta, ctypes.c_double), int(add_noise),
                            _p(K, ctypes.c_double))
    return K


def fit(Q, D, R, meta, x, y, theta, pi=PI_REF):
    """Returns (alpha, L row-major lower, logdet)."""
    meta, x, y = _prep(meta, x, y)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n = len(x)
    alpha = np.zeros(n)
    L = np.zeros((n, n))
    logdet = ctypes.c_double(0.0)
    rc = lib().medgp_oracle_fit(Q, D, R, ctypes.c_double(pi), n, _p(meta, ctypes.c_int32),
                                _p(x, ctypes.c_float), _p(y, ctypes.c_float),
                                _p(theta, ctypes.c_double), _p(alpha, ctypes.c_double),
                                _p(L, ctypes.c_double), ctypes.byref(logdet))
    if rc != 0:
        raise FloatingPointError("oracle: matrix not positive definite")
    return alpha, L, logdet.value


def predict(Q, D, R, meta, x, y, theta, meta_star, x_star, pi=PI_REF):
    """Returns (mean[m], var[m], status)."""
    meta, x, y = _prep(meta, x, y)
    meta_star, x_star, _ = _prep(meta_star, x_star)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    m = len(x_star)
    mean = np.zeros(m)
    var = np.zeros(m)
    status = ctypes.c_int(0)
    lib().medgp_oracle_predict(Q, D, R, ctypes.c_double(pi), len(x), _p(meta, ctypes.c_int32),
                               _p(x, ctypes.c_float), _p(y, ctypes.c_float),
                               _p(theta, ctypes.c_double), m, _p(meta_star, ctypes.c_int32),
                               _p(x_star, ctypes.c_float), _p(mean, ctypes.c_double),
                               _p(var, ctypes.c_double), ctypes.byref(status))
    return mean, var, status.value


def prior(ptype, xval, p0, p1, pi=PI_REF):
    lp = ctypes.c_double(0.0)
    dlp = ctypes.c_double(0.0)
    lib().medgp_oracle_prior(int(ptype), ctypes.c_double(xval), ctypes.c_float(p0),
                             ctypes.c_float(p1), ctypes.c_double(pi), ctypes.byref(lp),
                             ctypes.byref(dlp))
    return lp.value, dlp.value


# ---------------------------------------------------------------- compiled reference (oracle A)

def have_ref():
    return os.path.exists(REF_EVAL)


def write_case(path, Q, D, R, meta, x, y, theta, meta_star=None, x_star=None):
    """Flat text case file understood by oracle/ref/ref_eval.cpp."""
    with open(path, "w") as f:
        f.write(f"{Q} {D} {R} {len(x)} {len(theta)}\n")
        for m, a, b in zip(meta, x, y):
            f.write(f"{int(m)} {float(a):.9g} {float(b):.9g}\n")
        for t in theta:
            f.write(f"{float(t):.17g}\n")
        if meta_star is not None:
            f.write(f"{len(x_star)}\n")
            for m, a in zip(meta_star, x_star):
                f.write(f"{int(m)} {float(a):.9g}\n")


def ref_env():
    return dict(os.environ, OMP_NESTED="TRUE", OMP_MAX_ACTIVE_LEVELS="4")


def parse_ref_output(text):
    nums, secs = [], float("nan")
    for line in text.split("\n"):
        line = line.strip()
        if line.startswith("seconds_per_eval"):
            secs = float(line.split()[1])
            continue
        try:
            nums.append(float(line))
        except ValueError:
            pass  # the reference prints progress text to stdout
    return dict(ok=bool(int(nums[0])), nlml=nums[1], values=np.array(nums[2:]), seconds=secs)


def ref_eval(Q, D, R, meta, x, y, theta, mode=1, threads=1, repeat=1, meta_star=None,
             x_star=None):
    """Runs the compiled reference on one case.  Returns dict(ok, nlml, values, seconds)."""
    if not have_ref():
        raise FileNotFoundError(REF_EVAL)
    with tempfile.NamedTemporaryFile("w", suffix=".case", delete=False) as tf:
        path = tf.name
    try:
        write_case(path, Q, D, R, meta, x, y, theta, meta_star, x_star)
        out = subprocess.run([REF_EVAL, path, str(mode), str(threads), str(repeat)], check=True,
                             capture_output=True, text=True, env=ref_env()).stdout
    finally:
        os.unlink(path)
    return parse_ref_output(out)
