// FP64 micro-benchmark for B200: DFMA (vector pipe) vs DMMA.8x8x4 (tensor pipe) peak.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *out, int iters, double a, double b)
{
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double *out, int iters, double a, double b)
{
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double *out; cudaMalloc(&out, 148 * 8 * 1024 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2) {
        int threads = warps * 32, blocks = 148 * 2, iters = 20000;
        for (int which = 0; which < 2; which++) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                if (which == 0) k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
                else k_dmma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            double fl = which == 0 ? 2.0 * 16 * iters * (double)threads * blocks
                                   : 2.0 * 256 * 16 * iters * (double)warps * blocks;
            printf("%s warps/block=%2d blocks=%d : %.2f TFLOP/s (%.3f ms)\n", which ? "DMMA" : "DFMA", warps, blocks, fl / best / 1e9, best);
        }
    }
    return 0;
}
