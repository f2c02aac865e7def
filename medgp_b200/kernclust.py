"""Kernel clustering and population "mode kernel" estimation -- the step between training and
testing.  Host side in Python, as the reference's is, with the same
function names and file contract:

    medgpc/clustering/feature_extraction.py:62-98   extract_LMC_SM_feature, compute_sm_feature
    medgpc/clustering/cluster.py:6-46               run_clustering_top, run_sklearn_gmm
    medgpc/clustering/mode_estimate.py:242-450      output_mode_LMC_SM, compute_kde, compute_mode
    medgpc/clustering/kernclust.py:11-58            kernel_clustering_top
    medgpc/util/binaryIO.py:6-36                    read/write of raw-double files, read_train_kernel

What is B200-native here is the part that costs: the KDE mode of every hyper-parameter of every
kernel cluster -- D(D+1)/2 + 2 sets per cluster plus D noise levels, each O(n^2) in the number of
patients -- is ONE batched launch (medgp_cuda_kde_mode) instead of ~1500 statsmodels fits.  The
Gaussian-mixture clustering itself stays sklearn's (a library call in the reference too).  There is
no CPU implementation of the KDE in this module: without the CUDA library it raises.

Kept from the reference on purpose: the Python side uses the TRUE pi and treats v^2 as the spectral
variance (visualization/fastkernel.py:33-48) while the C++ side uses PI = 3.14159265 and v
(SURVEY.md appendix C.8); "lambda" <= 0 after the SVD refactorisation becomes 1e-15; plots are not
produced (visualization is out of scope)."""
from __future__ import annotations

import json
import os

import numpy as np


# ------------------------------------------------------------------ files (util/binaryIO.py)
def write_double_to_bin(filename, d_array):
    np.asarray(d_array, dtype=np.float64).tofile(filename)


def read_double_from_bin(filename):
    return np.fromfile(filename, dtype=np.float64)


def read_train_kernel(pan_array, kernel_dir):
    """(ids, theta rows) of the patients whose train_flag is 1 (binaryIO.py:20-36)."""
    valid_pan, valid_hyp = [], []
    for pan in pan_array:
        try:
            flag = np.atleast_1d(np.loadtxt(os.path.join(kernel_dir, f"train_flag_{pan}.txt"), dtype=int))[0]
            if flag:
                valid_hyp.append(read_double_from_bin(os.path.join(kernel_dir, f"train_hyp_{pan}.bin")))
                valid_pan.append(pan)
        except OSError:
            continue
    return np.asarray(valid_pan), np.asarray(valid_hyp)


# ------------------------------------------------------------------ features (feature_extraction.py)
def compute_B_matrix(Q, D, R, hyp):
    """B_q = A_q A_q^T + diag(kappa_q) from a flat theta (visualization/fastkernel.py:3-31)."""
    out = []
    for q in range(Q):
        A = np.asarray(hyp[D + q * D * R: D + (q + 1) * D * R], dtype=np.float64).reshape(D, R)
        lam = np.exp(np.asarray(hyp[D + Q * (D * R + 2) + q * D: D + Q * (D * R + 2) + (q + 1) * D], dtype=np.float64))
        out.append(A @ A.T + np.diag(lam))
    return out


def compute_sm_feature(mu, v):
    """72-point response of one spectral-mixture component on an hourly grid plus a periodicity
    flag (feature_extraction.py:87-98; compute_sm_1d / compute_k of fastkernel.py:33-48)."""
    x = np.arange(72, dtype=np.float64)
    rsq = np.clip(x * x, 0, np.inf)          # squared distance to the origin
    resp = np.exp(-2 * (np.pi ** 2) * rsq * v) * np.cos(2 * np.pi * np.sqrt(rsq) * mu)
    return np.hstack((resp, 10.0 if mu > np.pi * np.sqrt(v) else 0.0))


def extract_LMC_SM_feature(pan_array, hyp_array, Q, D, R):
    """One feature row per (patient, component) whose B_q is not numerically zero
    (feature_extraction.py:62-84)."""
    hyp_array = np.asarray(hyp_array, dtype=np.float64)
    assert hyp_array.shape[1] == D + Q * (D * R + 2 + D)
    comp_pan, comp_qidx, comp_feature = [], [], []
    for pan, hyp in zip(pan_array, hyp_array):
        B = compute_B_matrix(Q, D, R, hyp)
        for q in range(Q):
            if np.max(np.abs(B[q])) > 1e-10:
                mu = np.exp(hyp[D + Q * D * R + q])
                v = np.exp(2 * hyp[D + Q * (D * R + 1) + q])
                comp_pan.append(pan)
                comp_qidx.append(q)
                comp_feature.append(compute_sm_feature(mu, v))
    return np.asarray(comp_pan), np.asarray(comp_qidx), np.asarray(comp_feature)


# ------------------------------------------------------------------ clustering (cluster.py)
def run_sklearn_gmm(feature, max_cluster_num, init_num=50, max_iter_num=2000, random_state=None):
    from sklearn import mixture
    best = (np.inf, None, None)
    for n_components in range(1, max_cluster_num + 1):
        gmm = mixture.GaussianMixture(n_components=n_components, covariance_type="full", max_iter=max_iter_num,
                                      n_init=init_num, random_state=random_state)
        gmm.fit(feature)
        bic = gmm.bic(feature)
        print("BIC = {:.6f} for {} clusters".format(bic, n_components))
        if bic < best[0]:
            best = (bic, n_components, gmm.predict(feature))
    print("best cluster number using gmm clustering: {}".format(best[1]))
    return best[1], best[2]


def run_clustering_top(algorithm, feature, max_cluster_num=None, init_num=10, max_iter_num=2000, random_state=None):
    if max_cluster_num is None:
        max_cluster_num = 5
        print("Warning: maximum number of clusters not set; use default value {}".format(max_cluster_num))
    algorithm = str(algorithm)
    if algorithm == "None":
        print("Warning: clustering algorithm is not specified; skip clustering")
        return 1, np.zeros(feature.shape[0], dtype=int)
    if algorithm == "gmm":
        return run_sklearn_gmm(feature, max_cluster_num, init_num, max_iter_num, random_state)
    print("Error: not supported algorithm {}".format(algorithm))
    raise NotImplementedError


# ------------------------------------------------------------------ mode estimation (mode_estimate.py)
def bw_silverman(x):
    """Silverman's rule as the reference's KDE dependency applies it (statsmodels
    bandwidths.bw_silverman): 0.9 * min(std(ddof=1), IQR/1.349) * n**(-1/5)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    iqr = (np.percentile(x, 75) - np.percentile(x, 25)) / 1.349
    std = np.std(x, ddof=1)
    a = min(std, iqr) if iqr > 0 else std
    h = 0.9 * a * len(x) ** (-0.2)
    if h == 0:
        raise RuntimeError("Selected KDE bandwidth is 0. Cannot estimate density.")
    return h


def gpu_kde_modes(ctx, sets):
    """compute_kde + compute_mode(weighted=True) for every array in `sets`: one batched launch."""
    return ctx.kde_mode(sets, [bw_silverman(v) for v in sets])


def collect_mode_sets(pan_array, hyp_array, mixture_pan, mixture_index, mixture_cluster_num, mixture_cluster_assign, Q, D, R):
    """The 1-D samples whose KDE modes define the mode kernel, in a fixed order:
    D noise levels; then per cluster: mu, v, and the D(D+1)/2 upper-triangle entries of the
    per-patient B matrices (components of one patient that fall into the same cluster are added
    before the KDE, mode_estimate.py:352-376)."""
    pan_array = np.asarray(pan_array)
    hyp_array = np.asarray(hyp_array, dtype=np.float64)
    row_of = {p: i for i, p in enumerate(pan_array)}
    sets = [np.exp(hyp_array[:, d]) for d in range(D)]
    cluster_ids = np.unique(mixture_cluster_assign)
    assert len(cluster_ids) == mixture_cluster_num
    for cid in cluster_ids:
        idx = np.where(mixture_cluster_assign == cid)[0]
        rows = np.array([row_of[mixture_pan[c]] for c in idx])
        qq = np.asarray(mixture_index)[idx]
        sets.append(np.exp(hyp_array[rows, D + Q * D * R + qq]))           # mu
        sets.append(np.exp(hyp_array[rows, D + Q * D * R + Q + qq]))       # (square root of) v
        clust_pan = np.asarray(mixture_pan)[idx]
        all_B = []
        for pan in np.unique(clust_pan):
            hyp = hyp_array[row_of[pan]]
            Bq = compute_B_matrix(Q, D, R, hyp)
            B = np.zeros((D, D))
            for q in qq[clust_pan == pan]:
                B += Bq[q]
            all_B.append(B)
        all_B = np.asarray(all_B)
        for d1 in range(D):
            for d2 in range(d1, D):
                sets.append(all_B[:, d1, d2].copy())
    return sets


def assemble_mode_hyp(modes, newQ, D, R):
    """Mode theta from the KDE modes in collect_mode_sets order: noise logs, log mu, log v, and the
    mode B of every cluster refactored into A (first R left singular vectors scaled by the root
    singular values) and kappa = diag(B - A A^T), floored at 1e-15 (mode_estimate.py:403-418)."""
    modes = np.asarray(modes, dtype=np.float64)
    out = np.zeros(D + newQ * (D * R + 2 + D))
    out[:D] = np.log(modes[:D])
    pos = D
    npair = D * (D + 1) // 2
    for q in range(newQ):
        out[D + newQ * D * R + q] = np.log(modes[pos])
        out[D + newQ * (D * R + 1) + q] = np.log(modes[pos + 1])
        kde_B = np.zeros((D, D))
        k = pos + 2
        for d1 in range(D):
            for d2 in range(d1, D):
                kde_B[d1, d2] = kde_B[d2, d1] = modes[k]
                k += 1
        pos += 2 + npair
        U, S, _ = np.linalg.svd(kde_B)
        A_ = (U * np.sqrt(S))[:, 0:R]
        lam_ = np.diag(kde_B - A_ @ A_.T).copy()
        lam_[lam_ <= 0.0] = 1e-15
        out[D + newQ * (D * R + 2) + q * D: D + newQ * (D * R + 2) + (q + 1) * D] = np.log(lam_)
        out[D + q * D * R: D + (q + 1) * D * R] = A_.reshape(-1)
    return out


def output_mode_LMC_SM(fold, exp_param, pan_array, hyp_array, mixture_pan, mixture_index, mixture_cluster_num,
                       mixture_cluster_assign, kernclust_alg, ctx=None, kde_modes=None):
    """Estimates the mode kernel and writes <exp_kernel_dir>/{fold<f>|all}/<alg>_mode_mixture_num.txt and
    <alg>_mode_param.bin (mode_estimate.py:242-435).  ctx: a medgp_b200.api.Context (the KDE runs on
    its GPU).  kde_modes: tests inject the oracle here; the product path leaves it None."""
    Q, D, R = exp_param["Q"], exp_param["D"], exp_param["R"]
    sets = collect_mode_sets(pan_array, hyp_array, mixture_pan, mixture_index, mixture_cluster_num,
                             mixture_cluster_assign, Q, D, R)
    if kde_modes is not None:
        modes = kde_modes(sets)
    else:
        if ctx is None:
            raise RuntimeError("output_mode_LMC_SM needs a medgp_b200.api.Context: the KDE runs on the GPU, "
                               "there is no CPU fallback")
        modes = gpu_kde_modes(ctx, sets)
    mode_hyp = assemble_mode_hyp(modes, mixture_cluster_num, D, R)
    out_dir = os.path.join(exp_param["exp_kernel_dir"], "fold{}".format(fold) if fold != -1 else "all")
    os.makedirs(out_dir, exist_ok=True)
    np.savetxt(os.path.join(out_dir, "{}_mode_mixture_num.txt".format(kernclust_alg)), [mixture_cluster_num], fmt="%d")
    write_double_to_bin(os.path.join(out_dir, "{}_mode_param.bin".format(kernclust_alg)), mode_hyp)
    return mode_hyp


def kernel_clustering_top(exp_config, fold=-1, algorithm=None, ctx=None, kde_modes=None, random_state=None):
    """Train outputs -> component features -> clusters -> mode kernel files (kernclust.py:11-58)."""
    exp_param = json.load(open(exp_config, "r"))
    valid_pan = np.atleast_1d(np.genfromtxt(os.path.join(exp_param["data_dir"], exp_param["cohort_id_list"]), dtype=str))
    if fold != -1:
        cv_assign = np.atleast_1d(np.loadtxt(os.path.join(exp_param["cv_assign_file"]), dtype=int))
        valid_pan = valid_pan[np.where(cv_assign != fold)]
    kernel_pan, kernel_hyp = read_train_kernel(valid_pan, exp_param["exp_train_dir"])
    if exp_param["kernel"] != "LMC-SM":
        print("specified kernel type {} not supported".format(exp_param["kernel"]))
        raise NotImplementedError
    Q, D, R = exp_param["Q"], exp_param["D"], exp_param["R"]
    comp_pan, comp_qidx, comp_feature = extract_LMC_SM_feature(kernel_pan, kernel_hyp, Q, D, R)
    num, assign = run_clustering_top(algorithm, comp_feature, max_cluster_num=Q, random_state=random_state)
    return output_mode_LMC_SM(fold, exp_param, kernel_pan, kernel_hyp, comp_pan, comp_qidx, num, assign, str(algorithm),
                              ctx=ctx, kde_modes=kde_modes)
