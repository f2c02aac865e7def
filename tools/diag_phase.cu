// Cycle breakdown of the 64x64 diagonal-block routine (Cholesky factor + triangular inverse +
// forward-solve block) of linalg.cuh: one CTA, clock64 stamps at the phase boundaries.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o /tmp/diag_phase tools/diag_phase.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
__device__ long long g_clk[32];
#define MEDGP_PHASE(id) { if (threadIdx.x == 0) g_clk[id] = clock64(); }
#include "../medgp_b200/csrc/linalg.cuh"

__global__ void __launch_bounds__(MEDGP_DIAG_THREADS, 3) k_one(const EvalDesc *descs, int *fail)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(16) GjBufs gjb;
    __shared__ int s_fail;
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    diag_block_factor(descs[0], 0, smem, smem + kTileElems, &gjb, &s_fail, fail, false);
}

int main()
{
    const int n = 64;
    std::vector<double> A(kTileElems, 0.0), G(n * n);
    srand(1);
    for (auto &g : G) g = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int k = 0; k < n; k++) s += G[i * n + k] * G[j * n + k] / n;
            A[j * MEDGP_SLD + i] = s;
        }
    EvalDesc e = {};
    double *dM, *dX, *dXT, *drhs, *dblk;
    int *dfail;
    cudaMalloc(&dM, kTileElems * 8); cudaMalloc(&dX, kTileElems * 8); cudaMalloc(&dXT, kTileElems * 8);
    cudaMalloc(&drhs, 64 * 8); cudaMalloc(&dblk, 8); cudaMalloc(&dfail, 4);
    cudaMemset(drhs, 0, 64 * 8); cudaMemset(dfail, 0, 4);
    e.M = dM; e.dinv = dX; e.dinvT = dXT; e.rhs = drhs; e.blk = dblk; e.n = 64; e.npad = 64; e.T = 1; e.nrhs = 1;
    EvalDesc *dd;
    cudaMalloc(&dd, sizeof(e));
    cudaMemcpy(dd, &e, sizeof(e), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_one, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
    long long clk[32];
    for (int rep = 0; rep < 3; rep++) {
        cudaMemcpy(dM, A.data(), kTileElems * 8, cudaMemcpyHostToDevice);
        k_one<<<1, MEDGP_DIAG_THREADS, kGemmSmemBytes>>>(dd, dfail);
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(clk, g_clk, sizeof(clk));
    }
    const int ids[] = {0, 1, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    const char *names[] = {"entry", "tile loaded into smem", "J0 chol16+inverse16 (others: zero X)", "J0 rows below (DMMA)", "J0 next column block",
                           "J1 chol16+inv (others: deferred trailing)", "J1 rows below", "J1 next column block",
                           "J2 chol16+inv (others: deferred, X_10)", "J2 rows below", "J2 next column block",
                           "J3 chol16+inv (others: X_20, X_21)", "-", "X_3J level + logdet", "forward-solve block", "write-back"};
    for (int i = 1; i < 16; i++) printf("%-44s %7lld cycles\n", names[i], clk[ids[i]] - clk[ids[i - 1]]);
    printf("total %lld cycles = %.2f us at 1.965 GHz; status %s\n", clk[16] - clk[0], (clk[16] - clk[0]) / 1965.0, cudaGetErrorString(cudaGetLastError()));
    int f; cudaMemcpy(&f, dfail, 4, cudaMemcpyDeviceToHost);
    printf("fail flag %d\n", f);
    return 0;
}
