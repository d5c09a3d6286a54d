// Internal launch interface between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include "models.cuh"
#include "plant.cuh"

namespace dart {
struct LaunchInfo { int lanes, block_threads, grid, smem_bytes; };
int launch_solve(const KArgs& a, int lanes, int block_threads, cudaStream_t st, LaunchInfo* info);
int launch_pmpc_z(const KArgs& a, cudaStream_t st);
int launch_episode_pmpc(const KArgs& a, int T, const PlantArgs& plant, unsigned long long* counters, int lanes,
                        cudaStream_t st, LaunchInfo* info);
int launch_tilt_to_quat(int B, const double* u, double* q, cudaStream_t st);
}  // namespace dart
