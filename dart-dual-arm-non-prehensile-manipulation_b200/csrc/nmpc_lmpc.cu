// Instantiations of the NMPC solve kernel for the LMPC model (own translation unit: compiles in parallel).
#include "nmpc_kernel.cuh"

namespace dart {
int launch_solve_lmpc(const KArgs& a, int lanes, int block_threads, cudaStream_t st, LaunchInfo* info) {
    return launch_g<LmpcAxis>(a, lanes > 0 ? lanes : 32, block_threads, st, info);
}
}  // namespace dart
