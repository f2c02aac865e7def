"""ctypes binding of libmedgp_cuda.so (include/medgp_cuda.h) used by tests/ and bench.py.

This is plumbing only: every numerical result comes from the CUDA library.  There is no CPU
path -- loading fails loudly when the shared library is missing, and ``Context`` raises when
no sm_100 GPU is present.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MEDGP_LIB", os.path.join(HERE, "libmedgp_cuda.so"))  # MEDGP_LIB: experiments only
PI_REF = 3.14159265  # medgpc/src/util/global_settings.h:6
ORDER_FEATURE, ORDER_TIME, ORDER_GIVEN = 0, 1, 2  # include/medgp_cuda.h: MEDGP_ORDER_*

STAGES = ["prep", "assemble", "potrf", "diag", "solve", "trtri", "lauum", "grad", "predict"]

# every symbol include/medgp_cuda.h declares (tests check the library exports all of them)
SYMBOLS = [
    "medgp_cuda_create", "medgp_cuda_destroy", "medgp_cuda_last_error", "medgp_cuda_model",
    "medgp_cuda_num_hyp", "medgp_cuda_add_series", "medgp_cuda_add_series_ordered", "medgp_cuda_free_series",
    "medgp_cuda_clear_series", "medgp_cuda_nlml_grad", "medgp_cuda_nlml_grad_device",
    "medgp_cuda_sync", "medgp_cuda_predict", "medgp_cuda_predict_online", "medgp_cuda_debug_matrices", "medgp_cuda_debug_force_fail", "medgp_cuda_profile",
    "medgp_cuda_stage_times", "medgp_cuda_malloc", "medgp_cuda_free", "medgp_cuda_memcpy_h2d",
    "medgp_cuda_memcpy_d2h", "medgp_cuda_host_alloc", "medgp_cuda_host_free", "medgp_cuda_stream",
    "medgp_cuda_kde_mode", "medgp_cuda_export_factors", "medgp_cuda_add_series_batch", "medgp_cuda_free_series_batch", "medgp_cuda_scg_create", "medgp_cuda_scg_destroy", "medgp_cuda_scg_start", "medgp_cuda_scg_run",
    "medgp_cuda_scg_result", "medgp_cuda_scg_points", "medgp_cuda_scg_feed",
]


class StageTimes(ctypes.Structure):
    _fields_ = [("ms", ctypes.c_double * 9), ("launches", ctypes.c_longlong * 9),
                ("flops", ctypes.c_double * 9), ("bytes", ctypes.c_double * 9),
                ("evals", ctypes.c_longlong)]


class MedgpError(RuntimeError):
    pass


_lib = None


def build_library():
    """Compile the CUDA library in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], check=True)


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MEDGP_LIB", LIB_PATH)  # experiments: an alternative build of the same library
    if not os.path.exists(path):
        raise MedgpError(f"{path} is missing: build it with `make -C medgp_b200/csrc` "
                         "(there is no CPU fallback)")
    lib = ctypes.CDLL(path)
    vp, i, dp, ip = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.medgp_cuda_create.argtypes = [ctypes.POINTER(vp), i, ctypes.c_size_t]
    lib.medgp_cuda_destroy.argtypes = [vp]
    lib.medgp_cuda_destroy.restype = None
    lib.medgp_cuda_last_error.argtypes = [vp]
    lib.medgp_cuda_last_error.restype = ctypes.c_char_p
    lib.medgp_cuda_model.argtypes = [vp, i, i, i, ctypes.c_double]
    lib.medgp_cuda_num_hyp.argtypes = [vp]
    lib.medgp_cuda_add_series.argtypes = [vp, i, ip, fp, fp, ip]
    lib.medgp_cuda_add_series_ordered.argtypes = [vp, i, ip, fp, fp, i, ip]
    lib.medgp_cuda_predict_online.argtypes = [vp, i, ip, dp, dp, dp, ip]
    lib.medgp_cuda_free_series.argtypes = [vp, i]
    lib.medgp_cuda_add_series_batch.argtypes = [vp, i, ip, ip, fp, fp, i, ip]
    lib.medgp_cuda_free_series_batch.argtypes = [vp, i, ip]
    lib.medgp_cuda_clear_series.argtypes = [vp]
    lib.medgp_cuda_nlml_grad.argtypes = [vp, i, ip, dp, i, dp, dp, ip]
    lib.medgp_cuda_nlml_grad_device.argtypes = [vp, i, ip, vp, i, vp, vp, vp]
    lib.medgp_cuda_sync.argtypes = [vp]
    lib.medgp_cuda_predict.argtypes = [vp, i, ip, dp, ip, ip, fp, dp, dp, ip]
    lib.medgp_cuda_debug_matrices.argtypes = [vp, i, dp, dp, dp, dp, dp]
    lib.medgp_cuda_debug_force_fail.argtypes = [vp, i]
    lib.medgp_cuda_profile.argtypes = [vp, i]
    lib.medgp_cuda_stage_times.argtypes = [vp, ctypes.POINTER(StageTimes), i]
    lib.medgp_cuda_malloc.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.medgp_cuda_free.argtypes = [vp, vp]
    lib.medgp_cuda_memcpy_h2d.argtypes = [vp, vp, vp, ctypes.c_size_t]
    lib.medgp_cuda_memcpy_d2h.argtypes = [vp, vp, vp, ctypes.c_size_t]
    lib.medgp_cuda_host_alloc.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.medgp_cuda_host_free.argtypes = [vp, vp]
    lib.medgp_cuda_stream.argtypes = [vp]
    lib.medgp_cuda_stream.restype = vp
    bp = ctypes.POINTER(ctypes.c_byte)
    lib.medgp_cuda_kde_mode.argtypes = [vp, i, ip, dp, dp, dp, dp]
    lib.medgp_cuda_export_factors.argtypes = [vp, i, dp, fp, fp, dp, ip]
    lib.medgp_cuda_scg_create.argtypes = [vp, i, ctypes.POINTER(vp)]
    lib.medgp_cuda_scg_destroy.argtypes = [vp]
    lib.medgp_cuda_scg_destroy.restype = None
    lib.medgp_cuda_scg_start.argtypes = [vp, ip, dp, ip, bp, bp, fp]
    lib.medgp_cuda_scg_run.argtypes = [vp, i, ip]
    lib.medgp_cuda_scg_result.argtypes = [vp, dp, dp, ip]
    lib.medgp_cuda_scg_points.argtypes = [vp, dp, ip]
    lib.medgp_cuda_scg_feed.argtypes = [vp, dp, dp, ip]
    _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


class Context:
    """One GPU, one model shape (Q, D, R), many uploaded series."""

    def __init__(self, Q, D, R, device=0, workspace_bytes=0, pi=PI_REF):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        rc = self.lib.medgp_cuda_create(ctypes.byref(self.h), int(device), int(workspace_bytes))
        if rc != 0:
            self.h = None
            raise MedgpError(f"medgp_cuda_create failed with status {rc} "
                             "(no sm_100 GPU? there is no CPU fallback)")
        self.Q, self.D, self.R = Q, D, R
        self._pinned = []
        self._n = {}
        self._check(self.lib.medgp_cuda_model(self.h, Q, D, R, float(pi)))
        self.P = self.lib.medgp_cuda_num_hyp(self.h)

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.medgp_cuda_last_error(self.h)
            raise MedgpError(f"medgp_cuda status {rc}: {msg.decode() if msg else ''}")

    def close(self):
        if self.h:
            for p in self._pinned:  # arrays from pinned() must not be used after close()
                self.lib.medgp_cuda_host_free(self.h, p)
            self._pinned = []
            self.lib.medgp_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_model(self, Q, D, R, pi=PI_REF):
        """Change the kernel shape of this context (its series must have been cleared)."""
        self._check(self.lib.medgp_cuda_model(self.h, int(Q), int(D), int(R), float(pi)))
        self.Q, self.D, self.R = int(Q), int(D), int(R)
        self.P = self.lib.medgp_cuda_num_hyp(self.h)

    # ------------------------------------------------------------------ series
    def add_series(self, meta, x, y, order=ORDER_FEATURE):
        """order=ORDER_TIME uploads the series for predict_online (no gradients on it)."""
        meta = np.ascontiguousarray(meta, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        sid = ctypes.c_int(-1)
        self._check(self.lib.medgp_cuda_add_series_ordered(self.h, len(x), _ip(meta), _fp(x), _fp(y),
                                                           int(order), ctypes.byref(sid)))
        self._n[sid.value] = len(x)
        return sid.value

    def free_series(self, sid):
        self._check(self.lib.medgp_cuda_free_series(self.h, int(sid)))

    def clear_series(self):
        self._check(self.lib.medgp_cuda_clear_series(self.h))

    # ------------------------------------------------------------------ hot path, host buffers
    def pinned(self, shape, dtype=np.float64):
        """A numpy array backed by page-locked host memory (freed with the context)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape))
        p = ctypes.c_void_p()
        self._check(self.lib.medgp_cuda_host_alloc(self.h, n * dtype.itemsize, ctypes.byref(p)))
        self._pinned.append(p)
        buf = (ctypes.c_char * (n * dtype.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def nlml_grad(self, series_ids, theta, want_grad=True, out=None):
        """theta: (batch, P).  Returns (nlml[batch], grad[batch,P] or None, status[batch]).
        out = (nlml, grad, status) reuses caller arrays (page-locked ones are written by DMA)."""
        sids = np.ascontiguousarray(series_ids, dtype=np.int32)
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(len(sids), self.P)
        if out is not None:
            nlml, grad, status = out
        else:
            nlml = np.empty(len(sids))
            grad = np.empty((len(sids), self.P)) if want_grad else None
            status = np.empty(len(sids), dtype=np.int32)
        self._check(self.lib.medgp_cuda_nlml_grad(
            self.h, len(sids), _ip(sids), _dp(theta), int(want_grad), _dp(nlml),
            _dp(grad) if want_grad else None, _ip(status)))
        return nlml, grad, status

    def predict(self, series_ids, theta, star_offset, meta_star, x_star):
        sids = np.ascontiguousarray(series_ids, dtype=np.int32)
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(len(sids), self.P)
        star_offset = np.ascontiguousarray(star_offset, dtype=np.int32)
        meta_star = np.ascontiguousarray(meta_star, dtype=np.int32)
        x_star = np.ascontiguousarray(x_star, dtype=np.float32)
        m = int(star_offset[-1])
        mean, var = np.empty(m), np.empty(m)
        status = np.empty(len(sids), dtype=np.int32)
        self._check(self.lib.medgp_cuda_predict(
            self.h, len(sids), _ip(sids), _dp(theta), _ip(star_offset), _ip(meta_star),
            _fp(x_star), _dp(mean), _dp(var), _ip(status)))
        return mean, var, status

    def predict_online(self, series_ids, theta):
        """One-step-ahead imputation of every point of every (time-ordered) series with one
        factorisation each.  Returns (list of mean arrays, list of var arrays, status)."""
        sids = np.ascontiguousarray(series_ids, dtype=np.int32)
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(len(sids), self.P)
        ns = [self._n[int(s)] for s in sids]
        tot = int(sum(ns))
        mean, var = np.empty(tot), np.empty(tot)
        status = np.empty(len(sids), dtype=np.int32)
        self._check(self.lib.medgp_cuda_predict_online(self.h, len(sids), _ip(sids), _dp(theta),
                                                       _dp(mean), _dp(var), _ip(status)))
        cut = np.cumsum(ns)[:-1]
        return np.split(mean, cut), np.split(var, cut), status

    # ------------------------------------------------------------------ device-resident variant
    def malloc(self, nbytes):
        p = ctypes.c_void_p()
        self._check(self.lib.medgp_cuda_malloc(self.h, int(nbytes), ctypes.byref(p)))
        return p

    def free(self, p):
        self._check(self.lib.medgp_cuda_free(self.h, p))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self._check(self.lib.medgp_cuda_memcpy_h2d(self.h, dptr, arr.ctypes.data_as(ctypes.c_void_p), arr.nbytes))

    def d2h(self, arr, dptr):
        self._check(self.lib.medgp_cuda_memcpy_d2h(self.h, arr.ctypes.data_as(ctypes.c_void_p), dptr, arr.nbytes))

    def nlml_grad_device(self, series_ids, d_theta, want_grad, d_nlml, d_grad, d_status):
        sids = np.ascontiguousarray(series_ids, dtype=np.int32)
        self._check(self.lib.medgp_cuda_nlml_grad_device(
            self.h, len(sids), _ip(sids), d_theta, int(want_grad), d_nlml, d_grad, d_status))

    def sync(self):
        self._check(self.lib.medgp_cuda_sync(self.h))

    def stream(self):
        return self.lib.medgp_cuda_stream(self.h)

    # ------------------------------------------------------------------ taps and timings
    def debug_matrices(self, sid, theta, n, want=("K", "L", "alpha", "Kinv")):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        out = {}
        K = np.zeros((n, n)) if "K" in want else None
        L = np.zeros((n, n)) if "L" in want else None
        a = np.zeros(n) if "alpha" in want else None
        Ki = np.zeros((n, n)) if "Kinv" in want else None
        self._check(self.lib.medgp_cuda_debug_matrices(
            self.h, int(sid), _dp(theta), _dp(K) if K is not None else None,
            _dp(L) if L is not None else None, _dp(a) if a is not None else None,
            _dp(Ki) if Ki is not None else None))
        for k, v in (("K", K), ("L", L), ("alpha", a), ("Kinv", Ki)):
            if v is not None:
                out[k] = v
        return out

    def export_factors(self, sid, theta):
        """(alpha float32[n], L^-1 float32[n,n] lower, nlml, status) of a series uploaded with ORDER_GIVEN."""
        n = self._n[int(sid)]
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        alpha, linv = np.zeros(n, dtype=np.float32), np.zeros((n, n), dtype=np.float32)
        nlml, status = ctypes.c_double(0.0), ctypes.c_int(0)
        self._check(self.lib.medgp_cuda_export_factors(self.h, int(sid), _dp(theta), _fp(alpha), _fp(linv),
                                                       ctypes.byref(nlml), ctypes.byref(status)))
        return alpha, linv, nlml.value, status.value

    def kde_mode(self, sets, bandwidth, want_density=False):
        """Gaussian-KDE mode (density-weighted mean) of every 1-D array in `sets` (medgp_cuda_kde_mode)."""
        sets = [np.ascontiguousarray(v, dtype=np.float64).ravel() for v in sets]
        off = np.zeros(len(sets) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(v) for v in sets])
        data = np.concatenate(sets) if sets else np.zeros(0)
        bw = np.ascontiguousarray(bandwidth, dtype=np.float64)
        mode = np.empty(len(sets))
        dens = np.empty(len(data)) if want_density else None
        self._check(self.lib.medgp_cuda_kde_mode(self.h, len(sets), _ip(off), _dp(data), _dp(bw), _dp(mode),
                                                 _dp(dens) if want_density else None))
        return (mode, np.split(dens, off[1:-1])) if want_density else mode

    def scg_session(self, count):
        """Device-resident lock-step SCG over `count` instances (medgp_cuda_scg_*)."""
        return ScgSession(self, count)

    def force_fail(self, attempts):
        """Tests of the jitter path: the first `attempts` factorisation attempts count as failed."""
        self._check(self.lib.medgp_cuda_debug_force_fail(self.h, int(attempts)))

    def profile(self, enable=True):
        self._check(self.lib.medgp_cuda_profile(self.h, int(enable)))

    def stage_times(self, reset=False):
        st = StageTimes()
        self._check(self.lib.medgp_cuda_stage_times(self.h, ctypes.byref(st), int(reset)))
        return {name: dict(ms=st.ms[i], launches=st.launches[i], flops=st.flops[i], bytes=st.bytes[i])
                for i, name in enumerate(STAGES)} | {"evals": st.evals}


class ScgSession:
    """Python face of the device-resident optimiser session (include/medgp_cuda.h: medgp_cuda_scg_*)."""

    def __init__(self, ctx, count):
        self.ctx, self.lib, self.count, self.P = ctx, ctx.lib, int(count), ctx.P
        self.h = ctypes.c_void_p()
        ctx._check(self.lib.medgp_cuda_scg_create(ctx.h, self.count, ctypes.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.medgp_cuda_scg_destroy(self.h)
            self.h = ctypes.c_void_p()

    def start(self, series_ids, theta0, max_iteration, prior_type=None, prior_exp=None, prior_param=None):
        sids = np.ascontiguousarray(series_ids, dtype=np.int32)
        theta0 = np.ascontiguousarray(theta0, dtype=np.float64).reshape(self.count, self.P)
        budget = np.ascontiguousarray(np.broadcast_to(max_iteration, (self.count,)), dtype=np.int32)
        bp = ctypes.POINTER(ctypes.c_byte)
        if prior_type is None:
            pt = pe = pp = None
        else:
            prior_type = np.ascontiguousarray(prior_type, dtype=np.int8).reshape(self.count, self.P)
            prior_exp = np.ascontiguousarray(prior_exp, dtype=np.int8).reshape(self.count, self.P)
            prior_param = np.ascontiguousarray(prior_param, dtype=np.float32).reshape(self.count, self.P, 2)
            pt, pe, pp = prior_type.ctypes.data_as(bp), prior_exp.ctypes.data_as(bp), _fp(prior_param)
        self.ctx._check(self.lib.medgp_cuda_scg_start(self.h, _ip(sids), _dp(theta0), _ip(budget), pt, pe, pp))

    def run(self, super_steps):
        left = ctypes.c_int(0)
        self.ctx._check(self.lib.medgp_cuda_scg_run(self.h, int(super_steps), ctypes.byref(left)))
        return left.value

    def result(self):
        theta, loss = np.empty((self.count, self.P)), np.empty(self.count)
        evals = np.empty(self.count, dtype=np.int32)
        self.ctx._check(self.lib.medgp_cuda_scg_result(self.h, _dp(theta), _dp(loss), _ip(evals)))
        return theta, loss, evals

    def points(self):
        theta, wants = np.empty((self.count, self.P)), np.empty(self.count, dtype=np.int32)
        self.ctx._check(self.lib.medgp_cuda_scg_points(self.h, _dp(theta), _ip(wants)))
        return theta, wants

    def feed(self, f, grad, ok):
        f = np.ascontiguousarray(f, dtype=np.float64)
        grad = np.ascontiguousarray(grad, dtype=np.float64).reshape(self.count, self.P)
        ok = np.ascontiguousarray(ok, dtype=np.int32)
        self.ctx._check(self.lib.medgp_cuda_scg_feed(self.h, _dp(f), _dp(grad), _ip(ok)))
