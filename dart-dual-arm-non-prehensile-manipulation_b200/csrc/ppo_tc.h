// Launchers of the tensor-core layer-1 GEMMs of the PPO step (ppo_tc.cu), called from ppo.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dart_ppo_tc {
// Z[M,128] = X[gidx[m] or m, 0:520] W1^T (pre-activation, no bias); X 16-byte aligned
int l1_forward(int M, const float* X, const int64_t* gidx, const float* W1, float* Z, cudaStream_t st);
// number of minibatch splits (<= max_splits, each a whole number of 32-sample chunks, none empty) and samples per split
int l1_wgrad_splits(int M, int max_splits, int* per_out);
// split s: part[s * stride + n * 520 + f] = sum over its samples of dZ[m, n] X[.., f];  part[s * stride + 128 * 520 + n] = sum dZ[m, n]
int l1_wgrad(int M, int splits, int per, const float* dZ, const float* X, const int64_t* gidx, float* part, long stride,
             cudaStream_t st);
}  // namespace dart_ppo_tc
