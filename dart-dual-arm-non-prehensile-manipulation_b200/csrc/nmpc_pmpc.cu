// Instantiations of the NMPC solve kernel for the PMPC model (own translation unit: compiles in parallel).
#include "nmpc_kernel.cuh"

namespace dart {
int launch_solve_pmpc(const KArgs& a, int lanes, int block_threads, cudaStream_t st, LaunchInfo* info) {
    return launch_g<PmpcAxis>(a, lanes > 0 ? lanes : ((long)a.B * 2 > 8192 ? 8 : 16), block_threads, st, info);
}
int launch_episode_pmpc(const KArgs& a, int T, const PlantArgs& plant, unsigned long long* counters, int lanes,
                        cudaStream_t st, LaunchInfo* info) {
    EpisodeArgs ep{T, plant, counters};
    return launch_episode<PmpcAxis>(a, ep, lanes > 0 ? lanes : ((long)a.B * 2 > 8192 ? 8 : 16), st, info);
}
}  // namespace dart
