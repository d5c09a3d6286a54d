// Dispatcher plus the small epilogue kernels (PMPC z rows, tilt -> quaternion).
#include <cuda_runtime.h>

#include "models.cuh"
#include "launch.h"

namespace dart {

int launch_solve_pmpc(const KArgs& a, int lanes, int block_threads, cudaStream_t st, LaunchInfo* info);
int launch_solve_rmpc(const KArgs& a, int lanes, int block_threads, cudaStream_t st, LaunchInfo* info);
int launch_solve_lmpc(const KArgs& a, int lanes, int block_threads, cudaStream_t st, LaunchInfo* info);

int launch_solve(const KArgs& a, int lanes, int block_threads, cudaStream_t st, LaunchInfo* info) {
    switch (a.cfg.method) {
        case DART_PMPC: return launch_solve_pmpc(a, lanes, block_threads, st, info);
        case DART_RMPC: return launch_solve_rmpc(a, lanes, block_threads, st, info);
        case DART_LMPC: return launch_solve_lmpc(a, lanes, block_threads, st, info);
        default: return DART_ERR_ARG;
    }
}

// PMPC: z rows of the decision vector (one thread per instance; tiny, only when w_out is requested)
__global__ void pmpc_z_kernel(const KArgs a) {
    const int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst < a.B) pmpc_z_rollout(a, inst);
}

int launch_pmpc_z(const KArgs& a, cudaStream_t st) {
    if (!a.w_out) return DART_OK;
    pmpc_z_kernel<<<(a.B + 127) / 128, 128, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

// Euler xyz [u1, -u0, 0] -> wxyz (PMPC/main.py:107-116)
__global__ void tilt_to_quat_kernel(int B, const double* __restrict__ u, double* __restrict__ q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const double ax = u[2 * i + 1], ay = -u[2 * i];
    double sx, cx, sy, cy;
    sincos(0.5 * ax, &sx, &cx);
    sincos(0.5 * ay, &sy, &cy);
    const double cz = 1.0, sz = 0.0;
    q[4 * i + 0] = cx * cy * cz + sx * sy * sz;
    q[4 * i + 1] = sx * cy * cz - cx * sy * sz;
    q[4 * i + 2] = cx * sy * cz + sx * cy * sz;
    q[4 * i + 3] = cx * cy * sz - sx * sy * cz;
}

int launch_tilt_to_quat(int B, const double* u, double* q, cudaStream_t st) {
    if (B <= 0) return DART_OK;
    tilt_to_quat_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, u, q);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

}  // namespace dart
