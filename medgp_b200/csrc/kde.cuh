// kde.cuh -- batched kernel-density mode estimation for the population "mode kernel"
// (medgpc/clustering/mode_estimate.py:242-450).  Between training and
// testing the reference estimates, for every hyper-parameter of every kernel cluster, a Gaussian
// KDE over the cohort's fitted values (statsmodels KDEUnivariate, kernel "gau", Silverman
// bandwidth; mode_estimate.py:438-444) evaluated AT the data points, and takes the
// density-weighted mean of the data as the mode (compute_mode, weighted=True, :446-450):
//     dens_i = 1/(n h) sum_j phi((x_j - x_i) / h),   mode = sum_i x_i dens_i / sum_i dens_i
// That is O(n^2) per parameter and D(D+1)/2 + 2 parameters per cluster plus D noise levels:
// ~1500 independent sets of up to #patients values -- one launch here.
#pragma once
#include <cuda_runtime.h>

#define MEDGP_KDE_THREADS 256

// grid (ceil(n_max / 256), n_sets): thread = one evaluation point x_i of its set; the set's values
// stream through shared memory 256 at a time.
__global__ void __launch_bounds__(MEDGP_KDE_THREADS)
k_kde_density(const double *__restrict__ data, const int *__restrict__ offsets, const double *__restrict__ bw,
              double *__restrict__ dens)
{
    __shared__ double sx[MEDGP_KDE_THREADS];
    const int set = blockIdx.y, lo = offsets[set], n = offsets[set + 1] - lo;
    if ((int)(blockIdx.x * MEDGP_KDE_THREADS) >= n) return;
    const int i = blockIdx.x * MEDGP_KDE_THREADS + threadIdx.x;
    const double h = bw[set], inv_h = 1.0 / h;
    const double xi = i < n ? data[lo + i] : 0.0;
    double acc = 0.0;
    for (int j0 = 0; j0 < n; j0 += MEDGP_KDE_THREADS) {
        const int j = j0 + threadIdx.x;
        __syncthreads();
        sx[threadIdx.x] = j < n ? data[lo + j] : 0.0;
        __syncthreads();
        const int m = min(MEDGP_KDE_THREADS, n - j0);
        for (int t = 0; t < m; t++) {
            const double u = (sx[t] - xi) * inv_h;
            acc += exp(-0.5 * u * u);
        }
    }
    // 0.3989422804014327 = 1/sqrt(2 pi): the Gaussian kernel's normalisation
    if (i < n) dens[lo + i] = 0.3989422804014327 * acc / (h * (double)n);
}

// one CTA per set: mode = sum x dens / sum dens, fixed summation order
__global__ void __launch_bounds__(MEDGP_KDE_THREADS)
k_kde_mode(const double *__restrict__ data, const int *__restrict__ offsets, const double *__restrict__ dens,
           double *__restrict__ mode)
{
    __shared__ double s_num[MEDGP_KDE_THREADS], s_den[MEDGP_KDE_THREADS];
    const int set = blockIdx.x, lo = offsets[set], n = offsets[set + 1] - lo;
    double num = 0.0, den = 0.0;
    for (int i = threadIdx.x; i < n; i += MEDGP_KDE_THREADS) {
        num += data[lo + i] * dens[lo + i];
        den += dens[lo + i];
    }
    s_num[threadIdx.x] = num;
    s_den[threadIdx.x] = den;
    __syncthreads();
    for (int o = MEDGP_KDE_THREADS / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s_num[threadIdx.x] += s_num[threadIdx.x + o];
            s_den[threadIdx.x] += s_den[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) mode[set] = s_num[0] / s_den[0];
}
