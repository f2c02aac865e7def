"""Generates tests/golden/clustering.json by running the REFERENCE's own clustering code
(medgpc/clustering/feature_extraction.py, mode_estimate.py, imported from /root/reference) on
seeded synthetic "fitted" hyper-parameters.  Run here once; tests only read the committed JSON.

    python tests/golden/make_golden_clustering.py

The reference's Python side does not import in this image as it is: it needs matplotlib, seaborn
and statsmodels (absent) and uses names numpy 2 removed (np.float_, np.infty).  This script
  * aliases np.float_ / np.infty (no numerical effect),
  * stubs matplotlib / seaborn (plots are out of scope) and replaces the three plotting helpers
    mode_estimate imports by no-ops,
  * provides statsmodels.nonparametric.kde.KDEUnivariate from oracle/oracle_kde.py -- the numpy
    restatement of statsmodels' Gaussian KDE with Silverman bandwidth (NOT statsmodels itself:
    that one piece stays unpinned and is cross-checked against scipy in tests/test_kernclust.py).
Everything else -- which components count, the 72-point response features, which values enter
which KDE, the aggregation of B matrices per patient and cluster, the density-weighted mode, the
SVD refactorisation into A and kappa, the layout of the mode vector -- is the reference's code.
"""
import json
import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REFERENCE_ROOT = os.environ.get("MEDGP_REFERENCE", "/root/reference")

from medgp_b200 import synth  # noqa: E402
from oracle import oracle_kde  # noqa: E402


def import_reference_clustering():
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    if not hasattr(np, "infty"):
        np.infty = np.inf
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.backends",
                 "matplotlib.backends.backend_pdf", "matplotlib.patches", "matplotlib.colors", "matplotlib.cm",
                 "matplotlib.ticker", "mpl_toolkits", "mpl_toolkits.axes_grid1", "seaborn", "sklearn", "sklearn.mixture",
                 "pandas"):
        sys.modules.setdefault(name, mock.MagicMock())
    sm = types.ModuleType("statsmodels")
    smn = types.ModuleType("statsmodels.nonparametric")
    smk = types.ModuleType("statsmodels.nonparametric.kde")
    smk.KDEUnivariate = oracle_kde.KDEUnivariate
    sys.modules.update({"statsmodels": sm, "statsmodels.nonparametric": smn, "statsmodels.nonparametric.kde": smk})
    sys.path.insert(0, REFERENCE_ROOT)
    from medgpc.clustering import feature_extraction, mode_estimate
    for fn in ("plot_one_kernel", "plot_kde_hist", "plot_cluster_scatter_2d"):
        setattr(mode_estimate, fn, lambda *a, **k: None)
    return feature_extraction, mode_estimate


def fitted_thetas(Q, D, R, count, seed):
    """seeded stand-ins for fitted hyper-parameters: the reference's init distribution, with a few
    components switched off (A = 0, kappa tiny) so that not every (patient, component) counts"""
    th = synth.init_hyp_lmc_sm(Q, D, R, count, seed=seed)
    rng = np.random.default_rng(seed)
    for i in rng.choice(count, size=count // 5, replace=False):
        q = int(rng.integers(Q))
        th[i, D + q * D * R: D + (q + 1) * D * R] = 0.0
        th[i, D + Q * (D * R + 2) + q * D: D + Q * (D * R + 2) + (q + 1) * D] = np.log(1e-12)
    return th


def main():
    fe, me = import_reference_clustering()
    Q, D, R, count, seed = 2, 3, 2, 40, 91
    pans = np.array([f"p{i}" for i in range(count)])
    hyp = fitted_thetas(Q, D, R, count, seed)
    comp_pan, comp_qidx, comp_feature = fe.extract_kernel_feature("LMC-SM", Q, D, R, pans, hyp)
    gold = {"note": "outputs of the reference's medgpc/clustering code (KDE = oracle restatement of statsmodels)",
            "Q": Q, "D": D, "R": R, "count": count, "seed": seed,
            "comp_pan": comp_pan.tolist(), "comp_qidx": comp_qidx.tolist(),
            "feature_checksum": [float(comp_feature.sum()), float(np.abs(comp_feature).sum())],
            "feature_rows": {"0": comp_feature[0].tolist(), "last": comp_feature[-1].tolist()},
            "modes": []}
    with tempfile.TemporaryDirectory() as tmp:
        exp_param = {"kernel": "LMC-SM", "Q": Q, "D": D, "R": R, "exp_kernel_dir": os.path.join(tmp, "kernel"),
                     "exp_figure_dir": os.path.join(tmp, "figure")}
        assigns = {"one_cluster": np.zeros(len(comp_pan), dtype=int),
                   "two_clusters": (comp_feature[:, -1] > 5).astype(int) if len(np.unique(comp_feature[:, -1])) > 1
                   else (np.arange(len(comp_pan)) % 2)}
        for name, assign in assigns.items():
            num = len(np.unique(assign))
            mode = me.output_mode_kernel(fold=0, exp_param=exp_param, pan_array=pans, hyp_array=hyp, mixture_pan=comp_pan,
                                         mixture_index=comp_qidx, mixture_cluster_num=num, mixture_cluster_assign=assign,
                                         kernclust_alg=name, plotting_mode=1, plotting_param=None)
            written = np.fromfile(os.path.join(tmp, "kernel", "fold0", f"{name}_mode_param.bin"))
            qnum = int(open(os.path.join(tmp, "kernel", "fold0", f"{name}_mode_mixture_num.txt")).read())
            assert np.array_equal(written, mode) and qnum == num
            gold["modes"].append({"name": name, "assign": assign.tolist(), "cluster_num": num, "mode_hyp": mode.tolist()})
    with open(os.path.join(ROOT, "tests", "golden", "clustering.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote tests/golden/clustering.json:", len(comp_pan), "components,", [m["cluster_num"] for m in gold["modes"]])


if __name__ == "__main__":
    main()
