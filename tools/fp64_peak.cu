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
    // How many warps per SM sub-partition does the DMMA pipe need?  One CTA per SM, 1/2/3/4 warps per sub-partition
    // (16 independent accumulators per warp).
    for (int wps = 1; wps <= 4; wps++) {
        const int threads = 128 * wps, blocks = 148, iters = 20000;
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            k_dmma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("DMMA %d warp(s) per sub-partition: %.2f TFLOP/s\n", wps, 2.0 * 256 * 16 * iters * (double)(threads / 32) * blocks / best / 1e9);
    }
    // Do DFMA and DMMA run on the same datapath?  Both kernels at once (two streams, one 8-warp
    // CTA of each per SM): separate pipes would add up to about 71 TFLOP/s, one shared pipe stays at
    // about 35-37 in total.
    {
        cudaStream_t sa, sb;
        cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking);
        const int threads = 256, blocks = 148, iters = 40000;
        double *out2; cudaMalloc(&out2, 148 * 8 * 1024 * 8);
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaDeviceSynchronize();
            cudaEventRecord(e0, sa);
            cudaStreamWaitEvent(sb, e0, 0);
            k_dfma<<<blocks, threads, 0, sa>>>(out, iters, 1.0000001, 1e-9);
            k_dmma<<<blocks, threads, 0, sb>>>(out2, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1, sb);
            cudaStreamWaitEvent(sa, e1, 0);
            cudaEventRecord(e1, sa);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double f_dfma = 2.0 * 16 * iters * (double)threads * blocks, f_dmma = 2.0 * 256 * 16 * iters * 8.0 * blocks;
        printf("DFMA + DMMA concurrently (one 8-warp CTA of each per SM): %.3f ms, %.2f + %.2f = %.2f TFLOP/s in total\n", best,
               f_dfma / best / 1e9, f_dmma / best / 1e9, (f_dfma + f_dmma) / best / 1e9);
        for (int which = 0; which < 2; which++) {
            cudaDeviceSynchronize();
            cudaEventRecord(e0, sa);
            if (which == 0) k_dfma<<<blocks, threads, 0, sa>>>(out, iters, 1.0000001, 1e-9);
            else k_dmma<<<blocks, threads, 0, sa>>>(out2, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1, sa); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("  alone, same shape: %s %.3f ms = %.2f TFLOP/s\n", which ? "DMMA" : "DFMA", ms, (which ? f_dmma : f_dfma) / ms / 1e9);
        }
    }
    return 0;
}
