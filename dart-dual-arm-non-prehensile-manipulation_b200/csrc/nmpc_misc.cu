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

// Step hand-shake of the peer-rows gather (dart_peer_handshake): one thread per peer.  The rows of this step were stored by
// the solve kernel that precedes this launch on the stream, so they are complete; the system-scope fence orders them before
// the flag for an observer on another GPU.  Then every thread waits for ITS peer's flag in this GPU's own array.
struct PeerFlags { long long* f[DART_MAX_PEERS]; };

__global__ void peer_handshake_kernel(const PeerFlags pf, int n_peers, int my_rank, long long step, int* timed_out) {
    const int p = threadIdx.x;
    if (p >= n_peers) return;
    __threadfence_system();
    *reinterpret_cast<volatile long long*>(pf.f[p] + my_rank) = step;
    __threadfence_system();
    const volatile long long* mine = pf.f[my_rank] + p;
    const long long t0 = clock64();
    while (*mine < step) {
        __nanosleep(200);
        if (clock64() - t0 > 4000000000LL) {            // ~2 s at 1.9 GHz: a peer is gone
            if (timed_out) *timed_out = 1;
            break;
        }
    }
    __threadfence_system();
}

}  // namespace dart

extern "C" int dart_peer_handshake(int64_t* const* peer_flags, int32_t n_peers, int32_t my_rank, int64_t step, int32_t* timed_out,
                                   void* stream) {
    if (!peer_flags || n_peers < 1 || n_peers > DART_MAX_PEERS || my_rank < 0 || my_rank >= n_peers) return DART_ERR_ARG;
    dart::PeerFlags pf;
    for (int p = 0; p < DART_MAX_PEERS; ++p) {
        pf.f[p] = p < n_peers ? reinterpret_cast<long long*>(peer_flags[p]) : nullptr;
        if (p < n_peers && !pf.f[p]) return DART_ERR_ARG;
    }
    dart::peer_handshake_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pf, n_peers, my_rank, (long long)step, timed_out);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}
