// Dependent-chain latency probe (dev tool): cycles per op for one warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b) {
    __shared__ double sm[64];
    sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
    __syncwarp();
    double v = a; long long t0, t1; const int R = 4096;
    t0 = clock64(); for (int i = 0; i < R; ++i) v = fma(v, a, b); t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < R; ++i) v = 1.0 / (v + 1.5); t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < R; ++i) v = b / (v + 1.5); t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0);
    int idx = threadIdx.x;
    t0 = clock64(); for (int i = 0; i < R; ++i) { v += sm[idx]; idx = (idx + (int)v) & 31; } t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < R; ++i) v = __shfl_xor_sync(0xffffffffu, v, 1) + a; t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < R; ++i) v = sqrt(v * v + 1.0); t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < R; ++i) v = log(v + 2.0); t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0);
    float f = (float)v;
    t0 = clock64(); for (int i = 0; i < R; ++i) f = fmaf(f, (float)a, (float)b); t1 = clock64(); if (threadIdx.x == 0) cyc[7] = (t1 - t0);
    t0 = clock64(); for (int i = 0; i < R; ++i) { double s, c; sincos(v, &s, &c); v = s + c; } t1 = clock64(); if (threadIdx.x == 0) cyc[8] = (t1 - t0);
    out[threadIdx.x] = v + f;
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 256); cudaMalloc(&c, 128);
    k<<<1, 32>>>(o, c, 0.999, 0.001); k<<<1, 32>>>(o, c, 0.999, 0.001);
    long long h[9]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    const char* n[9] = {"dfma", "drcp(1/x)+add", "ddiv+add", "lds+add+idx", "shfl64+add", "dsqrt+fma", "dlog+add", "ffma", "dsincos+add"};
    for (int i = 0; i < 9; ++i) printf("%-16s %.1f cycles/iter\n", n[i], h[i] / 4096.0);
    return cudaGetLastError() != cudaSuccess;
}
