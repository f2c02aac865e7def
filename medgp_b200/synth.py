"""Deterministic synthetic patients and hyper-parameter draws.

The reference ships no data (MIMIC-III is access controlled), so every test and the
benchmark run on series produced here.  Layout follows the reference's loader
(medgpc/src/dataio/c_experiment.cpp:254-309): points are feature-major, ``meta[i]`` is the
position of the point's feature in the feature list, times are float32 hours rounded to six
decimals as the data files are (scripts/jmlr_mimic_heart_failure.py:285), values are float32
z-scores.  Hyper-parameters are drawn exactly as the reference's random initialisation does
(medgpc/src/dataio/c_experiment.cpp:418-441,493-564): glibc ``srand``/``rand() % 4096``.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

PI_REF = 3.14159265  # medgpc/src/util/global_settings.h:6

# scripts/opt_prior0.json:9-18 (identical in opt_prior2.json)
DEFAULT_BOUNDS = dict(noise=(0.15, 0.4), a=(-1.5, 1.5), period=(12.0, 72.0),
                      lengthscale=(6.0, 72.0), lam=(0.1, 0.5))


def hyp_counts(Q: int, D: int, R: int):
    """(n_lik, n_cov, P): medgpc/src/dataio/c_experiment.cpp:311-393."""
    n_cov = Q * (D * R + 2 + D)
    return D, n_cov, D + n_cov


def make_counts(D: int, n: int, rng: np.random.Generator) -> np.ndarray:
    """n points split over D features, at least 2 each (main_one_train.cpp:181-197)."""
    if n < 2 * D:
        raise ValueError("need n >= 2*D")
    extra = rng.multinomial(n - 2 * D, np.full(D, 1.0 / D))
    return (extra + 2).astype(np.int64)


def make_patient(D: int, n: int, seed: int, T: float = 240.0, counts=None):
    """Returns (meta int32[n], x float32[n], y float32[n]) feature-major."""
    rng = np.random.default_rng(seed)
    if counts is None:
        counts = make_counts(D, n, rng)
    meta, xs = [], []
    for d, c in enumerate(counts):
        t = np.sort(rng.uniform(0.5, T, size=int(c)))
        xs.append(np.round(t, 6).astype(np.float32))
        meta.append(np.full(int(c), d, dtype=np.int32))
    meta = np.concatenate(meta)
    x = np.concatenate(xs)
    y = rng.standard_normal(meta.shape[0]).astype(np.float32)
    return meta, x, y


def hyp_bounds(Q: int, D: int, R: int, bounds=None):
    """lb/ub per hyper-parameter in theta order, as medgpc/util/config.py:38-65 writes them."""
    b = dict(DEFAULT_BOUNDS)
    if bounds:
        b.update(bounds)
    seq = [b["noise"]] * D + [b["a"]] * (Q * D * R) + [b["period"]] * Q \
        + [b["lengthscale"]] * Q + [b["lam"]] * (Q * D)
    lb = np.array([s[0] for s in seq], dtype=np.float64)
    ub = np.array([s[1] for s in seq], dtype=np.float64)
    return lb, ub


class GlibcRand:
    """glibc srand()/rand() through libc, so draws are bit-identical to the reference's."""

    def __init__(self, seed: int):
        self._libc = ctypes.CDLL("libc.so.6")
        self._libc.srand(ctypes.c_uint(seed))

    def __call__(self) -> int:
        return int(self._libc.rand())


def _one_random(rnd, lb, ub, scale):
    # c_experiment.cpp:493-517 with flag_inv = flag_log = false
    temp = float(rnd() % 4096) + 1.0
    temp *= (ub - lb)
    temp = temp / 4096.0
    return scale * (temp + lb)


def init_hyp_lmc_sm(Q: int, D: int, R: int, num: int, seed: int = 718, bounds=None,
                    pi: float = PI_REF) -> np.ndarray:
    """``num`` random theta vectors, order and arithmetic of c_experiment.cpp:418-441,532-564."""
    lb, ub = hyp_bounds(Q, D, R, bounds)
    rnd = GlibcRand(seed)
    _, _, P = hyp_counts(Q, D, R)
    out = np.empty((num, P), dtype=np.float64)
    for k in range(num):
        for i in range(P):
            if i < D:
                v = math.log(_one_random(rnd, lb[i], ub[i], 1.0))
            elif i < D + Q * D * R:
                v = _one_random(rnd, lb[i], ub[i], 0.9 / math.sqrt(float(Q) * float(R)))
            elif i < D + Q * (D * R + 1):
                v = math.log(1.0 / _one_random(rnd, lb[i], ub[i], 1.0))
            elif i < D + Q * (D * R + 2):
                v = math.log(1.0 / (2 * pi * _one_random(rnd, lb[i], ub[i], 1.0)))
            else:
                v = math.log(_one_random(rnd, lb[i], ub[i], 0.1 / float(Q)))
            out[k, i] = v
    return out
