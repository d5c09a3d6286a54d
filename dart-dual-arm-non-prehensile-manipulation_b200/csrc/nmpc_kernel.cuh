// sm_100a kernel template: batched NMPC solve, one tile of G lanes of a warp per sub-problem.
#pragma once
#include <cuda_runtime.h>

#include "models.cuh"
#include "launch.h"
#include "plant.cuh"

namespace dart {

// LOCK = true: the tiles of a warp run in LOCKSTEP (same instruction stream, see the scan path of Solver::run), so every
// collective names the full warp with a compile-time mask and compiles to a bare SHFL / one WARPSYNC; the shuffle width
// still confines data to the tile.  LOCK = false: each tile names only its own lanes.  That mask is a lane-dependent
// value, so the compiler wraps every __shfl_sync / __syncwarp in a MATCH.ANY loop over the distinct masks of the warp
// (one pass per tile) and puts a WARPSYNC in front of every shuffle -- measured on the PMPC kernel: 49 such loops, and
// the collectives, not the arithmetic, set the time of the shuffle-heavy phases.
template <int G, bool LOCK = false>
struct DevTile {
    static constexpr int kLanes = G;
    static constexpr bool kLockstep = LOCK || G == 32;
    unsigned mask_;
    int ln;
    __device__ __forceinline__ DevTile() {
        const int l = threadIdx.x & 31;
        ln = l & (G - 1);
        mask_ = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (l - ln));
    }
    __device__ __forceinline__ unsigned m() const { return kLockstep ? 0xffffffffu : mask_; }
    __device__ __forceinline__ unsigned warp_ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    __device__ __forceinline__ bool group_all(bool p) const { return __all_sync(m(), p) != 0; }
    __device__ __forceinline__ int lane() const { return ln; }
    __device__ __forceinline__ int size() const { return G; }
    __device__ __forceinline__ void sync() const { __syncwarp(m()); }
    __device__ __forceinline__ void block_sync() const { __syncthreads(); }
    __device__ __forceinline__ bool block_any(bool p) const { return __syncthreads_or(p ? 1 : 0) != 0; }
    __device__ __forceinline__ double shfl(double v, int src) const { return __shfl_sync(m(), v, src, G); }
    __device__ __forceinline__ double sum(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m(), v, o, G);
        return v;
    }
    __device__ __forceinline__ double max(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(m(), v, o, G));
        return v;
    }
    // max of values that are >= 0 (|.|, or maxima seeded with 0): for those the IEEE-754 order is the unsigned integer
    // order of the bit pattern, so two REDUX instructions replace log2(G) shuffle levels; a NaN still wins the max
    // (full-warp tiles only: with two sub-warp masks in one warp REDUX measured slower than the shuffle tree)
    __device__ __forceinline__ double max_nonneg(double v) const {
        if (G != 32) return max(v);
        const unsigned hi = (unsigned)__double2hiint(v) & 0x7fffffffu, lo = (unsigned)__double2loint(v);
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
        return __hiloint2double((int)mh, (int)ml);
    }
    // min of positive values (products z * slack of interior iterates), same integer-order argument
    __device__ __forceinline__ double min_pos(double v) const {
        if (G != 32) return min(v);
        const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
        const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
        return __hiloint2double((int)mh, (int)ml);
    }
    __device__ __forceinline__ double min(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(m(), v, o, G));
        return v;
    }
};

// lockstep tiles: the scan path of Solver::run (2-state / 1-input axis problems, one lane per stage + the terminal cost)
template <class M, int G, int NC>
struct TileFor {
    static constexpr bool scan = M::SERIAL_RICCATI && M::NX == 2 && M::NU == 1 && NC > 0 && (G == NC + 1 || 2 * G == NC + 1);
    static constexpr bool lock = G < 32 && (scan || !M::SERIAL_RICCATI);
    using type = DevTile<G, lock>;
};

// models whose axis problems run on full-warp tiles of their own (one warp per axis): see solve_block
template <class M, int G>
constexpr bool kAxisHandover = M::NAXIS > 1 && G == 32 && !M::SERIAL_RICCATI;

// M::MIN_BLOCKS: occupancy hint (blocks of M::MAX_THREADS per SM) that caps registers where shared memory leaves room.
// One batch solve by this block: every tile solves its sub-problem, then the axes of an instance are combined.
template <class M, int G, int NC>
__device__ __forceinline__ void solve_block(const KArgs& a, const int ws_stride, double* smem) {
    constexpr int NAX = M::NAXIS;
    using TileT = typename TileFor<M, G, NC>::type;
    const TileT tile;
    const int tpb = blockDim.x / G;
    const int tib = threadIdx.x / G;
    const long prob = (long)blockIdx.x * tpb + tib;
    const int inst = (int)(prob / NAX), axis = (int)(prob % NAX);
    double* slot = smem + (size_t)tib * ws_stride;
    const bool active = inst < a.B;
    double J = 0.0, kkt = 0.0;
    int32_t status = ST_CONVERGED, iters = 0;
    const BlockCtx bc{smem, ws_stride, tpb, (int)threadIdx.x};
    if (TileT::kLockstep && G < 32) {
        // lockstep tiles execute one instruction stream per warp: a tile past the end of the batch solves the last
        // instance again (same data, same trajectory) and writes nothing
        solve_one<M, TileT, NC>(tile, a, active ? inst : a.B - 1, axis, true, active, bc, slot + kSlot, J, status, iters, kkt);
    } else {
        solve_one<M, TileT, NC>(tile, a, inst, axis, active, active, bc, slot + kSlot, J, status, iters, kkt);
    }
    if (NAX == 1) {
        if (active && tile.lane() == 0) {
            a.J[inst] = J;
            if (a.status) a.status[inst] = status;
            if (a.iters) a.iters[inst] = iters;
            store_result_row(a, inst, a.u0[(long)inst * 2], a.u0[(long)inst * 2 + 1], J, (double)status);
        }
        return;
    }
    if constexpr (kAxisHandover<M, G>) {
        // Axes solved by DIFFERENT warps (full-warp tiles): no warp waits for its sibling.  Each axis publishes its record
        // and bumps the instance's arrival counter; whichever axis arrives last combines the records in axis order (so
        // the result does not depend on who that was), writes the outputs and re-arms the counter.  A warp that is done
        // leaves -- with one warp per block its registers and shared memory go to the next block straight away
        // (measured before: 14 % of the LMPC kernel's warp samples sat at the block barrier that used to be here).
        if (a.axis_sync != nullptr) {
            if (active && tile.lane() == 0) {
                double* part = a.axis_part + ((long)inst * NAX + axis) * 4;
                __stcg(part + 0, J); __stcg(part + 1, (double)status); __stcg(part + 2, (double)iters); __stcg(part + 3, kkt);
                __threadfence();
                const int arrived = atomicAdd(a.axis_sync + inst, 1);
                if (arrived == NAX - 1) {
                    __threadfence();
                    double Js = 0.0;
                    int32_t st = 0, itx = 0;
                    for (int ax = 0; ax < NAX; ++ax) {
                        const double* s2 = a.axis_part + ((long)inst * NAX + ax) * 4;
                        Js += __ldcg(s2 + 0);
                        const int32_t sx = (int32_t)__ldcg(s2 + 1);
                        if (status_rank(sx) > status_rank(st)) st = sx;
                        itx = max(itx, (int32_t)__ldcg(s2 + 2));
                    }
                    a.J[inst] = Js;
                    if (a.status) a.status[inst] = st;
                    if (a.iters) a.iters[inst] = itx;
                    if (a.rows || a.n_peers > 0)
                        store_result_row(a, inst, __ldcg(a.u0 + (long)inst * 2), __ldcg(a.u0 + (long)inst * 2 + 1), Js, (double)st);
                    a.axis_sync[inst] = 0;
                }
            }
            return;
        }
    }
    // combine the axes of one instance (adjacent tiles of the block), in a fixed order.  All tiles of the block leave
    // the solve loop together (block-uniform exit), so a block barrier orders the slot writes before the reads.
    if (tile.lane() == 0) {
        slot[0] = J;
        slot[1] = (double)status;
        slot[2] = (double)iters;
        slot[3] = kkt;
    }
    __syncthreads();
    if (active && axis == 0 && tile.lane() == 0) {
        double Js = 0.0;
        int32_t st = 0, itx = 0;
        for (int ax = 0; ax < NAX; ++ax) {
            const double* s2 = slot + (size_t)ax * ws_stride;
            Js += s2[0];
            if (status_rank((int32_t)s2[1]) > status_rank(st)) st = (int32_t)s2[1];
            itx = max(itx, (int32_t)s2[2]);
        }
        a.J[inst] = Js;
        if (a.status) a.status[inst] = st;
        if (a.iters) a.iters[inst] = itx;
        // u0 of every axis was stored before the barrier above
        if (a.rows || a.n_peers > 0) store_result_row(a, inst, a.u0[(long)inst * 2], a.u0[(long)inst * 2 + 1], Js, (double)st);
    }
}

template <class M, int G, int NC>
__global__ void __launch_bounds__(M::MAX_THREADS, M::MIN_BLOCKS) nmpc_solve_kernel(const KArgs a, const int ws_stride) {
    extern __shared__ __align__(16) double smem[];
    solve_block<M, G, NC>(a, ws_stride, smem);
}

// Persistent closed-loop episode: T times { solve the block's instances, advance their surrogate plant,
// update the episode metrics } in ONE launch.  Instances are independent, so blocks never synchronise with each other;
// inside a block the state written by the plant thread is visible to the solving tiles after the block barrier.
struct EpisodeArgs {
    int T;
    PlantArgs plant;                       // plant.u = a.u0, plant.state = a.x0, plant.target = a.ref
    unsigned long long* counters;          // [0] += iterations of every solve, [1] += solves that did not converge
};

template <class M, int G, int NC>
__global__ void __launch_bounds__(M::MAX_THREADS, M::MIN_BLOCKS) nmpc_episode_kernel(const KArgs a, const int ws_stride,
                                                                                     const EpisodeArgs e) {
    extern __shared__ __align__(16) double smem[];
    const int ipb = (blockDim.x / G) / M::NAXIS;                 // instances of this block
    const int inst = blockIdx.x * ipb + (int)threadIdx.x;        // the plant thread's instance
    const bool plant_thread = (int)threadIdx.x < ipb && inst < a.B;
    for (int step = 0; step < e.T; ++step) {
        solve_block<M, G, NC>(a, ws_stride, smem);
        __syncthreads();                                         // u0 / status / iters of this block's instances are written
        if (plant_thread) {
            atomicAdd(&e.counters[0], (unsigned long long)a.iters[inst]);
            if (a.status[inst] != ST_CONVERGED) atomicAdd(&e.counters[1], 1ull);
            plant_step_one(e.plant, inst);
        }
        __syncthreads();                                         // the new states are visible to the solving tiles
    }
}

template <class M, int G, int NC>
static int launch_n(const KArgs& a, int block_threads, cudaStream_t st, LaunchInfo* info, const EpisodeArgs* ep = nullptr) {
    const int ws = Workspace<M>::doubles(a.N) + kSlot;
    // Shared memory has 16 banks of 8 bytes.  In the serial sweeps all lanes of a tile read the same address and the
    // 32/G tiles of a warp differ by the workspace stride, so the stride is padded to make them land on distinct
    // banks (stride = 16/T mod 16 for T tiles per warp); a stride that is a multiple of 16 would serialise them.
    constexpr int T = 32 / G;
    constexpr int want = (T > 1) ? (16 / T) % 16 : 0;
    int ws_stride = ws;
    while (T > 1 && (ws_stride % 16) != want) ++ws_stride;
    // the vectorised sweeps (Workspace::kVec) use 128-bit shared-memory accesses: every problem's workspace must start on a
    // 16-byte boundary (`want` is even for the tile widths those models run with; kSlot and doubles() are even)
    if (Workspace<M>::kVec && (ws_stride & 1)) ++ws_stride;
    int bt = block_threads > 0 ? block_threads : 0;
    // lanes that must share a block: the axes of an instance, unless they hand their results over through global memory
    const bool handover = kAxisHandover<M, G> && a.axis_sync != nullptr && ep == nullptr;
    const int unit = handover ? G : G * M::NAXIS;
    // device limits and the kernel's opt-in shared-memory size are cached PER DEVICE (a process may hold handles on
    // several GPUs): per-launch driver queries cost microseconds
    constexpr int kMaxDev = 64;
    static int max_smem_dev[kMaxDev] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return DART_ERR_CUDA;
    if (max_smem_dev[dev] == 0) cudaDeviceGetAttribute(&max_smem_dev[dev], cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int max_smem = max_smem_dev[dev];
    if (bt == 0) {
        // measured on B200 (tools/sweep.py): few problems -> 4 warps per block; a filled GPU -> one warp per
        // block, so shared memory (the occupancy limiter) packs at warp granularity.
        const long probs = (long)a.B * M::NAXIS;
        bt = (probs * G > 148L * 32 * 16) ? (handover ? 32 : M::BT_LARGE) : 128;
        if (bt < unit) bt = unit;
    }
    if (bt > M::MAX_THREADS) bt = M::MAX_THREADS;
    // block size: whole warps AND whole instances (the axes of an instance are combined through the shared-memory slots
    // of ONE block, solve_block: slot + ax * ws_stride), i.e. a multiple of lcm(32, G * NAXIS)
    int step = 32;
    while (step % unit != 0) step += 32;
    if (bt % step != 0 || bt < unit) return DART_ERR_ARG;
    int tpb = bt / G;
#ifdef DART_SMEM_PAD
    ws_stride += DART_SMEM_PAD / 8;       // occupancy experiment: waste shared memory per problem
#endif
    size_t smem = (size_t)tpb * ws_stride * sizeof(double);
    while (smem > (size_t)max_smem && bt > step) { bt -= step; tpb = bt / G; smem = (size_t)tpb * ws_stride * sizeof(double); }
    if (smem > (size_t)max_smem || bt < unit || (!handover && tpb % M::NAXIS != 0)) return DART_ERR_UNSUPPORTED;
    const long probs = (long)a.B * M::NAXIS;
    const int grid = (int)((probs + tpb - 1) / tpb);
    if (ep) {
        auto kern = nmpc_episode_kernel<M, G, NC>;
        static size_t smem_set_ep[kMaxDev] = {0};
        if (smem > smem_set_ep[dev]) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DART_ERR_CUDA;
            smem_set_ep[dev] = smem;
        }
        kern<<<grid, bt, smem, st>>>(a, ws_stride, *ep);
    } else {
        auto kern = nmpc_solve_kernel<M, G, NC>;
        static size_t smem_set[kMaxDev] = {0};          // per instantiation and device
        if (smem > smem_set[dev]) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DART_ERR_CUDA;
            smem_set[dev] = smem;
        }
        kern<<<grid, bt, smem, st>>>(a, ws_stride);
    }
    if (info) { info->lanes = G; info->block_threads = bt; info->grid = grid; info->smem_bytes = (int)smem; }
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

template <class M, int G>
static int launch_t(const KArgs& a, int block_threads, cudaStream_t st, LaunchInfo* info) {
    if (a.N == M::NDEF) return launch_n<M, G, M::NDEF>(a, block_threads, st, info);
    return launch_n<M, G, 0>(a, block_threads, st, info);
}

// Persistent episode: the reference horizon and the two tile widths the batch launcher picks for PMPC.
template <class M>
static int launch_episode(const KArgs& a, const EpisodeArgs& ep, int lanes, cudaStream_t st, LaunchInfo* info) {
    if (a.N != M::NDEF) return DART_ERR_UNSUPPORTED;
    if (lanes == 8) return launch_n<M, 8, M::NDEF>(a, 0, st, info, &ep);
    if (lanes == 16) return launch_n<M, 16, M::NDEF>(a, 0, st, info, &ep);
    return DART_ERR_UNSUPPORTED;
}

template <class M>
static int launch_g(const KArgs& a, int lanes, int block_threads, cudaStream_t st, LaunchInfo* info) {
    switch (lanes) {
        case 2: if constexpr (M::NX <= 2) return launch_t<M, 2>(a, block_threads, st, info); return DART_ERR_ARG;
        case 4: if constexpr (M::NX <= 5) return launch_t<M, 4>(a, block_threads, st, info); return DART_ERR_ARG;
        case 8: return launch_t<M, 8>(a, block_threads, st, info);
        case 16: return launch_t<M, 16>(a, block_threads, st, info);
        case 32: return launch_t<M, 32>(a, block_threads, st, info);
        default: return DART_ERR_ARG;
    }
}


}  // namespace dart
