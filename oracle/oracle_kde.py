"""TEST ORACLE for the mode-kernel KDE: numpy restatement of what
medgpc/clustering/mode_estimate.py:438-450 asks of its dependency.

    compute_kde(data, test_x):  KDEUnivariate(data).fit(kernel="gau", bw="silverman").evaluate(test_x)
    compute_mode(data, density, weighted=True):  nansum(data * density) / nansum(density)

The algorithm lives in a third-party package that is NOT in /root/reference and not installed in
this image: statsmodels (un-pinned in the reference's setup.py; any release of the 0.9-0.14
line behaves the same here).  Restated from its published source:
  * statsmodels/nonparametric/bandwidths.py  bw_silverman(x) = 0.9 * A * n**(-1/5),
      A = min(std(x, ddof=1), IQR / 1.349) (IQR from scoreatpercentile 25/75, linear
      interpolation; A = std when the IQR is 0); a zero bandwidth raises RuntimeError
  * statsmodels/nonparametric/kde.py  KDEUnivariate.evaluate(point) = kernel.density(endog, point)
  * statsmodels/sandbox/nonparametric/kernels.py  CustomKernel.density(xs, x) =
      1/(h n) * sum_j K((xs_j - x) / h),  Gaussian K(u) = 0.3989422804014327 * exp(-u**2 / 2)
PARITY UNPINNED against statsmodels itself (absent here); the density formula is cross-checked
against scipy.stats.gaussian_kde with the same bandwidth (tests/test_kernclust.py), and everything
around the KDE is pinned to the reference's own code (tests/golden/make_golden_clustering.py).
Imported only by tests/ and tests/golden/."""
import numpy as np


def bw_silverman(x):
    x = np.asarray(x, dtype=np.float64).ravel()
    iqr = (np.percentile(x, 75) - np.percentile(x, 25)) / 1.349
    std = np.std(x, ddof=1)
    a = min(std, iqr) if iqr > 0 else std
    h = 0.9 * a * len(x) ** (-0.2)
    if h == 0:
        raise RuntimeError("Selected KDE bandwidth is 0. Cannot estimate density.")
    return h


def kde_density(data, test_x, h=None):
    data = np.asarray(data, dtype=np.float64).ravel()
    test_x = np.asarray(test_x, dtype=np.float64).ravel()
    h = bw_silverman(data) if h is None else h
    u = (data[:, None] - test_x[None, :]) / h
    return 0.3989422804014327 * np.exp(-u ** 2 / 2.0).sum(axis=0) / (h * len(data))


def kde_mode(data):
    data = np.asarray(data, dtype=np.float64).ravel()
    dens = kde_density(data, data)
    return np.nansum(data * dens) / np.nansum(dens)


class KDEUnivariate:
    """Stand-in with the statsmodels call surface mode_estimate.py uses; lets the golden script run
    the REFERENCE's own mode estimation code on top of this restatement."""

    def __init__(self, endog):
        self.endog = np.asarray(endog, dtype=np.float64).ravel()
        self.bw = None

    def fit(self, kernel="gau", bw="silverman", **kw):
        assert kernel == "gau" and bw == "silverman"
        self.bw = bw_silverman(self.endog)
        return self

    def evaluate(self, point):
        return kde_density(self.endog, point, self.bw)
