// Internal launch interface between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include "models.cuh"

namespace dart {
struct LaunchInfo { int lanes, block_threads, grid, smem_bytes; };
int launch_solve(const KArgs& a, int lanes, int block_threads, cudaStream_t st, LaunchInfo* info);
int launch_pmpc_z(const KArgs& a, cudaStream_t st);
int launch_tilt_to_quat(int B, const double* u, double* q, cudaStream_t st);
}  // namespace dart
