// Accuracy of exp_nonpos (common.cuh) against expl() on the host: max and mean error in ulps over
// uniform and log-uniform samples of x in [-708, 0], checked and unchecked variants.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o build_exp/exp_ulp tools/exp_ulp.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../medgp_b200/csrc/common.cuh"

template <bool CHECKED>
__global__ void k_exp(const double *x, double *y, int n)
{
    __shared__ double s_tab[MEDGP_EXP_TAB];
    exp_tab_stage(s_tab);
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double xa[1] = {x[i]};
        double out[1];
        exp_nonpos<1, CHECKED>(xa, out, s_tab);
        y[i] = out[0];
    }
}

int main()
{
    const int n = 20000000;
    std::vector<double> x(n), y(n);
    srand48(7);
    for (int i = 0; i < n; i++) {
        const double u = drand48();
        x[i] = (i & 1) ? -708.0 * u : -exp(log(1e-8) + u * (log(708.0) - log(1e-8)));  // uniform / log-uniform
    }
    double *dx, *dy;
    cudaMalloc(&dx, n * 8); cudaMalloc(&dy, n * 8);
    cudaMemcpy(dx, x.data(), n * 8, cudaMemcpyHostToDevice);
    for (int variant = 0; variant < 2; variant++) {
        if (variant == 0) k_exp<true><<<1184, 256>>>(dx, dy, n);
        else k_exp<false><<<1184, 256>>>(dx, dy, n);
        cudaMemcpy(y.data(), dy, n * 8, cudaMemcpyDeviceToHost);
        double worst = 0.0, sum = 0.0, worst_x = 0.0;
        for (int i = 0; i < n; i++) {
            const long double ref = expl((long double)x[i]);
            const double rd = (double)ref;
            const double ulp = nextafter(rd, INFINITY) - rd;
            const double err = (double)fabsl(((long double)y[i] - ref) / (long double)ulp);
            sum += err;
            if (err > worst) { worst = err; worst_x = x[i]; }
        }
        printf("%s: max error %.3f ulp (at x = %.17g), mean %.3f ulp, %d samples, status %s\n", variant ? "unchecked" : "checked", worst,
               worst_x, sum / n, n, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
