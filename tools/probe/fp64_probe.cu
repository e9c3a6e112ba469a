// fp64_probe.cu -- measures FP64 latency / throughput on the device (used to size the FP64-heavy kernels)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double *out, int n, long long *cyc)
{
    double a = out[0], b = out[1];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = __fma_rn(a, b, 1.0);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[2] = a;
}
__global__ void lat_div(double *out, int n, long long *cyc)
{
    double a = out[0], b = out[1];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = __ddiv_rn(a, b) + 1.0;
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[1] = t1 - t0;
    out[3] = a;
}
__global__ void lat_rcp(double *out, int n, long long *cyc)
{
    double a = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = __drcp_rn(a) + 1.0;
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[2] = t1 - t0;
    out[4] = a;
}
__global__ void thr(double *out, int n)
{
    double a0 = out[0] + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7, b = out[1];
    for (int i = 0; i < n; ++i) {
        a0 = __fma_rn(a0, b, 1.0); a1 = __fma_rn(a1, b, 1.0); a2 = __fma_rn(a2, b, 1.0); a3 = __fma_rn(a3, b, 1.0);
        a4 = __fma_rn(a4, b, 1.0); a5 = __fma_rn(a5, b, 1.0); a6 = __fma_rn(a6, b, 1.0); a7 = __fma_rn(a7, b, 1.0);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[5] = a0;
}
#include "../../aruco_slam_b200/csrc/core.h"
__global__ void otsu_probe(const int *hist, long long *cyc, int *thr_out)
{
    __shared__ double q1s[256], ys[256], mu1s[256];
    __shared__ int h[256];
    const int lane = threadIdx.x;
    for (int i = lane; i < 256; i += 32) h[i] = hist[i];
    __syncwarp();
    int n = 1024;
    int lo = 256, hi = -1;
    for (int i = 0; i < 256; ++i) if (h[i]) { if (lo == 256) lo = i; hi = i; }
    for (int i = lane; i < 256; i += 32) { double p, ip; b2a::otsu_bin_inputs(i, h[i], n, p, ip); q1s[i] = p; mu1s[i] = ip; ys[i] = -1.0; }
    __syncwarp();
    long long t0 = clock64();
    if (lane == 0) b2a::otsu_prefix(lo, hi, q1s);
    __syncwarp();
    long long t1 = clock64();
    for (int i = lane; i < 256; i += 32) if (i >= lo && i <= hi) ys[i] = b2a::otsu_bin(q1s[i]);
    __syncwarp();
    long long t2 = clock64();
    if (lane == 0) b2a::otsu_chain(lo, hi, q1s, ys, mu1s);
    __syncwarp();
    long long t3 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = hi - lo + 1; thr_out[0] = (int)(mu1s[hi] * 1000); }
}
int main()
{
    {
        int hh[256]; for (int i = 0; i < 256; ++i) hh[i] = 0;
        for (int i = 20; i < 240; ++i) hh[i] = (i % 7 == 0) ? 3 : ((i < 60 || i > 200) ? 8 : 1);
        int *dh; long long *dc; int *dt;
        cudaMalloc(&dh, 1024); cudaMalloc(&dc, 64); cudaMalloc(&dt, 16);
        cudaMemcpy(dh, hh, 1024, cudaMemcpyHostToDevice);
        otsu_probe<<<1, 32>>>(dh, dc, dt);
        long long hc[4]; cudaMemcpy(hc, dc, 32, cudaMemcpyDeviceToHost);
        printf("otsu alone: bins %lld, prefix %lld cycles, per-bin %lld, chain %lld cycles\n", hc[3], hc[0], hc[1], hc[2]);
    }
    double *d; long long *c;
    cudaMalloc(&d, 64); cudaMalloc(&c, 64);
    double h[8] = {1.0000001, 0.9999999, 0, 0, 0, 0, 0, 0};
    cudaMemcpy(d, h, 64, cudaMemcpyHostToDevice);
    const int n = 4096;
    for (int threads : {1, 32}) {
        lat<<<1, threads>>>(d, n, c); lat_div<<<1, threads>>>(d, n, c); lat_rcp<<<1, threads>>>(d, n, c);
        long long hc[3];
        cudaMemcpy(hc, c, 24, cudaMemcpyDeviceToHost);
        printf("threads %2d: dependent DFMA %.1f cycles, DDIV+DADD %.1f cycles, DRCP+DADD %.1f cycles\n", threads, (double)hc[0] / n, (double)hc[1] / n, (double)hc[2] / n);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps : {1, 4, 8, 16, 32}) {
        const int iters = 20000;
        thr<<<148, warps * 32>>>(d, 100);
        cudaEventRecord(e0);
        thr<<<148, warps * 32>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fma = 148.0 * warps * 32 * 8.0 * iters;
        printf("throughput %2d warps/SM: %.2f TFMA/s = %.1f FP64 TFLOP/s, %.2f lanes/clk/SM at 1.9 GHz\n", warps, fma / ms / 1e9, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.9e9);
    }
    return 0;
}
