// FP64 FMA-pipe peak microbenchmark: the roofline denominator for the solver kernels
// (MEASURED_PEAKS.json has HBM and bf16 tensor peaks only).  16 independent DFMA chains per thread,
// full occupancy, timed with CUDA events.
#include <cuda_runtime.h>
#include "../../include/dart_b200.h"

__global__ void __launch_bounds__(256) fp64_fma_kernel(double* out, int iters, double a, double b) {
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (double)(threadIdx.x + i) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 123.456) out[0] = s;   // keep the chains alive without a store on the hot path
}

extern "C" int dart_measure_fp64_tflops(int device, double* tflops) {
    if (!tflops) return DART_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return DART_ERR_NO_DEVICE; }
    if (cudaSetDevice(device) != cudaSuccess) return DART_ERR_CUDA;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    double* d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return DART_ERR_ALLOC;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = sms * 8, threads = 256;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fp64_fma_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return DART_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return DART_OK;
}
