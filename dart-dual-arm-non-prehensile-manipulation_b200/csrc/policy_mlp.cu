// LMPC parameter-adaptation policy on the device (the only dense contraction of the hot path).
//
//   policy_mlp_kernel   Policy.mean_net forward (LMPC/src/controller/rlmpc2.py:33-46,71-80):
//                       [B,520] f32 -> Linear(520,64) -> tanh -> Linear(64,64) -> tanh -> Linear(64,34).
//                       Two persistent CTAs per SM walk 128-row tiles.  Layer 1 streams the observation tile and
//                       W1 through a 2-stage TMA/mbarrier pipeline (x2 CTAs, so one CTA streams while the other is in
//                       its epilogue) into tcgen05.mma (kind::tf32, M=128, N=64,
//                       fp32 accumulate in TMEM).  The epilogue warps read the accumulator with tcgen05.ld, add
//                       the bias, apply tanh and write the activations straight back to shared memory in the
//                       K-major 128B-swizzled operand layout, so layers 2 and 3 run as further tcgen05.mma on
//                       chip (W2, W3 resident in shared memory); only the [B,34] means go back to HBM.
//                       HBM traffic = 2080 B in + 136 B out per instance (the algorithmic minimum).
//   policy_obs_kernel   observation build of rlmpc2.py:641-668: base vector [state, target, control, current_k]
//                       (float32-rounded), Welford running mean/variance, normalise, append to the 10-deep history.
//   policy_param_kernel logit-space parameter update + smoothed, soft-clipped write-back (rlmpc2.py:742-759, 606-616).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <new>

#include "../../include/dart_b200.h"
#include "tc_common.cuh"

namespace {

constexpr int OBS = 520, HID = 64, ACT = 34, N3 = 48;
constexpr int MT = 128;                 // rows (instances) per tile = UMMA M
constexpr int BK = 32;                  // fp32 elements per 128-byte swizzle row
constexpr int NKB = (OBS + BK - 1) / BK;   // 17 K blocks, the last one zero-filled by TMA past column 520
constexpr int A_BYTES = MT * 128, B_BYTES = HID * 128, W3_BYTES = N3 * 128;
// shared-memory map for a pipeline of S stages: [A stages | W1 stages | W2 | W3 | H | barriers]
template <int S> struct Smem {
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + S * A_BYTES;
    static constexpr int OFF_W2 = OFF_B + S * B_BYTES;
    static constexpr int OFF_W3 = OFF_W2 + 2 * B_BYTES;
    static constexpr int OFF_H = OFF_W3 + 2 * W3_BYTES;
    static constexpr int OFF_BAR = OFF_H + 2 * A_BYTES;
    static constexpr int BYTES = OFF_BAR + 256 + 1024;       // + alignment slack
};
constexpr int NTHREADS = 192;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t ACC1 = 0, ACC2 = 64, ACC3 = 128;

using namespace dart_tc;

struct MlpArgs {
    int B, ntiles;
    const float *b1, *b2, *b3;
    float* mean;
};

// epilogue helper: 64 accumulator columns -> +bias -> tanh -> K-major SW128 operand tile in shared memory
__device__ __forceinline__ void acc_to_hidden(uint32_t tmem_acc, int row, const float* __restrict__ bias, uint8_t* hbuf) {
#pragma unroll
    for (int c0 = 0; c0 < HID; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_acc + c0, r);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int col = c0 + 4 * q;
            float4 v;
            v.x = tanhf(__uint_as_float(r[4 * q + 0]) + __ldg(bias + col + 0));
            v.y = tanhf(__uint_as_float(r[4 * q + 1]) + __ldg(bias + col + 1));
            v.z = tanhf(__uint_as_float(r[4 * q + 2]) + __ldg(bias + col + 2));
            v.w = tanhf(__uint_as_float(r[4 * q + 3]) + __ldg(bias + col + 3));
            const int kblk = col >> 5, chunk = (col & 31) >> 2;
            uint8_t* dst = hbuf + kblk * A_BYTES + row * 128 + ((chunk ^ (row & 7)) << 4);
            *reinterpret_cast<float4*>(dst) = v;
        }
    }
}

// STAGES = 2: two CTAs per SM (one streams while the other is in its epilogue) -- large batches.
// STAGES = 4: one CTA per SM with a deeper pipeline -- batches of at most one tile per SM, where latency rules.
template <int STAGES>
__global__ void __launch_bounds__(NTHREADS, STAGES == 2 ? 2 : 1)
policy_mlp_kernel(const __grid_constant__ CUtensorMap tm_obs, const __grid_constant__ CUtensorMap tm_w1,
                  const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_w3, const MlpArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int OFF_A = Smem<STAGES>::OFF_A, OFF_B = Smem<STAGES>::OFF_B, OFF_W2 = Smem<STAGES>::OFF_W2,
                  OFF_W3 = Smem<STAGES>::OFF_W3, OFF_H = Smem<STAGES>::OFF_H, OFF_BAR = Smem<STAGES>::OFF_BAR;
    // barriers: full[S], empty[S], wbar, acc_full[3], h_ready ; then the TMEM base slot
    const uint32_t bar0 = sbase + OFF_BAR;
    auto FULL = [&](int s) { return bar0 + 8 * s; };
    auto EMPTY = [&](int s) { return bar0 + 8 * (STAGES + s); };
    const uint32_t WBAR = bar0 + 8 * (2 * STAGES);
    auto ACCF = [&](int l) { return bar0 + 8 * (2 * STAGES + 1 + l); };
    const uint32_t HRDY = bar0 + 8 * (2 * STAGES + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * (2 * STAGES + 5));

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        mbar_init(WBAR, 1);
        for (int l = 0; l < 3; ++l) mbar_init(ACCF(l), 1);
        mbar_init(HRDY, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(WBAR, 2 * B_BYTES + 2 * W3_BYTES);
            tma_load_2d(sbase + OFF_W2, &tm_w2, 0, 0, WBAR);
            tma_load_2d(sbase + OFF_W2 + B_BYTES, &tm_w2, BK, 0, WBAR);
            tma_load_2d(sbase + OFF_W3, &tm_w3, 0, 0, WBAR);
            tma_load_2d(sbase + OFF_W3 + W3_BYTES, &tm_w3, BK, 0, WBAR);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                for (int kb = 0; kb < NKB; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(EMPTY(s), ph ^ 1);
                    mbar_expect_tx(FULL(s), A_BYTES + B_BYTES);
                    tma_load_2d(sbase + OFF_A + s * A_BYTES, &tm_obs, kb * BK, tile * MT, FULL(s));
                    tma_load_2d(sbase + OFF_B + s * B_BYTES, &tm_w1, kb * BK, 0, FULL(s));
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t ID64 = idesc_tf32(MT, HID), ID48 = idesc_tf32(MT, N3);
            mbar_wait(WBAR, 0);
            uint32_t it = 0, hph = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                for (int kb = 0; kb < NKB; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(FULL(s), ph);
                    fence_after();
                    const uint64_t da = sdesc(sbase + OFF_A + s * A_BYTES), db = sdesc(sbase + OFF_B + s * B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) umma_tf32(tmem + ACC1, da + 2 * k, db + 2 * k, ID64, (kb | k) != 0);
                    umma_commit(EMPTY(s));
                }
                umma_commit(ACCF(0));
                // layer 2: h1 (shared) x W2
                mbar_wait(HRDY, hph); hph ^= 1;
                fence_after();
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint64_t da = sdesc(sbase + OFF_H + kk * A_BYTES), db = sdesc(sbase + OFF_W2 + kk * B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) umma_tf32(tmem + ACC2, da + 2 * k, db + 2 * k, ID64, (kk | k) != 0);
                }
                umma_commit(ACCF(1));
                // layer 3: h2 (shared) x W3 (34 rows, zero-filled to 48)
                mbar_wait(HRDY, hph); hph ^= 1;
                fence_after();
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint64_t da = sdesc(sbase + OFF_H + kk * A_BYTES), db = sdesc(sbase + OFF_W3 + kk * W3_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) umma_tf32(tmem + ACC3, da + 2 * k, db + 2 * k, ID48, (kk | k) != 0);
                }
                umma_commit(ACCF(2));
            }
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        uint8_t* hbuf = smem + OFF_H;
        uint32_t tph = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, tph ^= 1) {
            mbar_wait(ACCF(0), tph);
            fence_after();
            acc_to_hidden(tl + ACC1, row, a.b1, hbuf);
            fence_before();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(HRDY);
            mbar_wait(ACCF(1), tph);
            fence_after();
            acc_to_hidden(tl + ACC2, row, a.b2, hbuf);
            fence_before();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(HRDY);
            mbar_wait(ACCF(2), tph);
            fence_after();
            // layer-3 result: stage the [128,34] tile in shared memory (the activation buffer is free once ACC3 is
            // complete) and write it out with coalesced 16-byte stores -- the tile is contiguous in global memory.
            float* stage = reinterpret_cast<float*>(hbuf);
#pragma unroll
            for (int c0 = 0; c0 < N3; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tl + ACC3 + c0, r);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < ACT) stage[row * ACT + c0 + j] = __uint_as_float(r[j]) + __ldg(a.b3 + c0 + j);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            {
                const long row0 = (long)tile * MT;
                const int valid = (a.B - row0 < MT) ? (int)(a.B - row0) : MT;
                const int nflt = valid * ACT;
                float* g = a.mean + row0 * ACT;                   // 16-byte aligned: MT * ACT * 4 is a multiple of 16
                const int et = (warp - 2) * 32 + lane;
                const int nv = nflt >> 2;
                for (int i = et; i < nv; i += 128) reinterpret_cast<float4*>(g)[i] = reinterpret_cast<const float4*>(stage)[i];
                for (int i = (nv << 2) + et; i < nflt; i += 128) g[i] = stage[i];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");        // the next tile's epilogue reuses the buffer
            fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// ======================================================================================= FP32-fidelity forward
// policy_mlp3_kernel: the production forward.  The reference's Policy runs in FP32 (torch never enables TF32), and a
// single TF32 pass (policy_mlp_kernel above) is three decimal digits short of that, so:
//   layer 1 (K = 520, 84 % of the flops and ALL of the HBM traffic) stays on tcgen05 but as a 3xTF32 product,
//       A W = A_hi W_hi + A_hi W_lo + A_lo W_hi      (hi = top 19 bits, lo = fp32(x - hi); dropped term ~ 2^-22),
//     W_hi / W_lo are split once on the host; A arrives by TMA as fp32 and serves as A_hi as it is (the tensor core ignores
//     the low 13 mantissa bits of a TF32 operand), four "split" warps write A_lo = A - top19(A) into a second buffer
//     (generic -> async proxy fence), then one thread issues the 8 MMAs of the chunk; fp32 accumulation in TMEM.
//   layers 2 and 3 (64x64 and 64x34 per row) run in plain FP32 FMAs on the CUDA cores, register-tiled, weights and
//     activations in shared memory -- the reference's arithmetic up to summation order, and no hi/lo operand copies of
//     W2, W3 and of the activations in shared memory.
// One persistent CTA per SM, warp-specialised: warp 0 TMA producer of the observation chunks (5-stage ring, 16 KB each),
// warp 14 TMA producer of the [W1_hi ; W1_lo] chunks (3-stage ring, 16 KB each, L2-resident source), warps 2-5 split
// (A_lo double buffer), warp 1 MMA issuer, warps 6-13 epilogue.  The layer-1 accumulator is double-buffered in TMEM, so the
// stream of tile t+1 runs under the epilogue (layers 2, 3, store) of tile t.
namespace v3 {
constexpr int NA = 5;                                                   // A ring: fp32 observation chunks landing from HBM (16 KB each)
constexpr int NW = 3;                                                   // W ring: [W1_hi chunk ; W1_lo chunk] from L2 (16 KB each)
constexpr int NL = 2;                                                   // A_lo double buffer (written by the split warps)
constexpr int OFF_A = 0;
constexpr int OFF_W = OFF_A + NA * A_BYTES;
constexpr int OFF_L = OFF_W + NW * 2 * B_BYTES;
// layers 2-3 operand block, one bulk copy from the blob dart_policy_create prepares: W2 and W3 transposed [k][j], biases
constexpr int W3LD = 40;                                                // 34 action columns padded to 40
constexpr int OFF_W2T = OFF_L + NL * A_BYTES;                           // [64 k][64 j] fp32
constexpr int OFF_W3T = OFF_W2T + HID * HID * 4;                        // [64 k][40 j] fp32
constexpr int OFF_BIAS = OFF_W3T + HID * W3LD * 4;                      // b1[64] b2[64] b3[40]
constexpr int BLOB_BYTES = HID * HID * 4 + HID * W3LD * 4 + (HID + HID + W3LD) * 4;
constexpr int OFF_HS = OFF_BIAS + (HID + HID + W3LD) * 4;               // activations [k][row] fp32; reused as the [128, 34] result tile
constexpr int OFF_BAR = OFF_HS + HID * MT * 4;
constexpr int NBAR = 2 * NA + 2 * NW + 2 * NL + 5;                      // fullA emptyA fullW emptyW split lfree accf[2] acce[2] blob
constexpr int BYTES = OFF_BAR + 8 * NBAR + 16 + 1024;                   // + tmem slot + alignment slack
static_assert(BLOB_BYTES % 16 == 0 && OFF_W2T % 16 == 0, "bulk copy granularity");
constexpr int NEPI = 256;                                               // epilogue threads (warps 6..13)
constexpr int NTHR = 64 + 128 + NEPI + 32;                              // + warp 14: W producer
constexpr uint32_t TCOLS = 256;                                         // two 128-column layer-1 accumulators ([.. W_hi | .. W_lo] partial sums)
static_assert(MT * ACT <= HID * MT, "the result tile fits in the activation buffer");
static_assert(BYTES <= 232448, "shared memory budget of one CTA per SM");
}  // namespace v3

__device__ __forceinline__ unsigned long long pack2(float x, float y) {
    unsigned long long v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x), "f"(y));
    return v;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
// two FP32 FMAs in one instruction (FFMA2, sm_100): acc += a * b, element-wise on register pairs
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
// tanh(x) = 1 - 2 / (1 + e^(2x)) with the SFU exponential and reciprocal (each ~1e-7 relative): absolute error <= 3e-7,
// the same size as the rounding of an FP32 tanh; |x| large saturates cleanly (e^(2x) -> 0 or inf)
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = exp2f(x * 2.885390081777927f);            // e^(2x); exp2f compiles to MUFU.EX2 (+ denormal scaling)
    return 1.0f - __fdividef(2.0f, 1.0f + e);
}

struct Mlp3Args {
    int B, ntiles;
    const float* blob;                          // [W2 transposed [k][j] | W3 transposed, 40 columns | b1 b2 b3]: v3::BLOB_BYTES, built at create
    float* mean;
};

__global__ void __launch_bounds__(v3::NTHR, 1)
policy_mlp3_kernel(const __grid_constant__ CUtensorMap tm_obs, const __grid_constant__ CUtensorMap tm_w1h,
                   const __grid_constant__ CUtensorMap tm_w1l, const Mlp3Args a) {
    using namespace v3;
    extern __shared__ uint8_t smem_raw[];
#ifdef DART_MLP3_CLOCK
    const long long t_entry = clock64();
#endif
    // 1024-byte alignment (SWIZZLE_128B atoms) by an OFFSET into the __shared__ array: a pointer rebuilt from an integer
    // would lose its address space and every access below would compile to a generic LD/ST instead of LDS/STS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = sbase + OFF_BAR;
    auto FULLA = [&](int s) { return bar0 + 8 * s; };
    auto EMPTYA = [&](int s) { return bar0 + 8 * (NA + s); };
    auto FULLW = [&](int s) { return bar0 + 8 * (2 * NA + s); };
    auto EMPTYW = [&](int s) { return bar0 + 8 * (2 * NA + NW + s); };
    auto SPLIT = [&](int l) { return bar0 + 8 * (2 * NA + 2 * NW + l); };
    auto LFREE = [&](int l) { return bar0 + 8 * (2 * NA + 2 * NW + NL + l); };
    auto ACCF = [&](int b) { return bar0 + 8 * (2 * NA + 2 * NW + 2 * NL + b); };
    auto ACCE = [&](int b) { return bar0 + 8 * (2 * NA + 2 * NW + 2 * NL + 2 + b); };
    const uint32_t BLOB = bar0 + 8 * (2 * NA + 2 * NW + 2 * NL + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * NBAR);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(FULLA(s), 1); mbar_init(EMPTYA(s), 1); }
        for (int s = 0; s < NW; ++s) { mbar_init(FULLW(s), 1); mbar_init(EMPTYW(s), 1); }
        for (int l = 0; l < NL; ++l) { mbar_init(SPLIT(l), 128); mbar_init(LFREE(l), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(ACCF(b), 1); mbar_init(ACCE(b), v3::NEPI); }
        mbar_init(BLOB, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;

    // The observation chunks (HBM, ~1 us away) and the weight chunks (L2) ride separate rings with their own producers, and
    // the A ring is the deep one: a stage is busy from the TMA issue until the MMAs that read it retire, so with S stages in
    // flight an SM streams S x 16 KB per (memory latency + split + MMA) -- at S = 3, with the weights in the same stage, that
    // was 18 GB/s per SM (measured: 15 of the kernel's 26 us at 16 384 rows), a third of the SM's share of HBM.
    if (warp == 0) {
        // ===== TMA producer, observation chunks =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                for (int kb = 0; kb < NKB; ++kb, ++it) {
                    const int s = it % NA;
                    mbar_wait(EMPTYA(s), ((it / NA) & 1) ^ 1);
                    mbar_expect_tx(FULLA(s), A_BYTES);
                    tma_load_2d(sbase + OFF_A + s * A_BYTES, &tm_obs, kb * BK, tile * MT, FULLA(s));
                }
            }
        }
    } else if (warp == 14) {
        // ===== TMA producer, W1 hi / lo chunks (and, first, the layers 2-3 operand block) =====
        if (lane == 0) {
            mbar_expect_tx(BLOB, BLOB_BYTES);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sbase + OFF_W2T), "l"(a.blob), "r"(BLOB_BYTES), "r"(BLOB) : "memory");
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                for (int kb = 0; kb < NKB; ++kb, ++it) {
                    const int s = it % NW;
                    mbar_wait(EMPTYW(s), ((it / NW) & 1) ^ 1);
                    mbar_expect_tx(FULLW(s), 2 * B_BYTES);
                    tma_load_2d(sbase + OFF_W + s * 2 * B_BYTES, &tm_w1h, kb * BK, 0, FULLW(s));
                    tma_load_2d(sbase + OFF_W + s * 2 * B_BYTES + B_BYTES, &tm_w1l, kb * BK, 0, FULLW(s));
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t ID64 = idesc_tf32(MT, HID), ID128 = idesc_tf32(MT, 2 * HID);
            uint32_t it = 0;
            int lt = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++lt) {
                const int b = lt & 1;
                mbar_wait(ACCE(b), (((uint32_t)lt >> 1) & 1) ^ 1);            // the epilogue has drained this accumulator
                fence_after();
                const uint32_t acc = tmem + (uint32_t)(b * 2 * HID);
                for (int kb = 0; kb < NKB; ++kb, ++it) {
                    const int sa = it % NA, sw = it % NW, l = it % NL;
                    mbar_wait(FULLW(sw), (it / NW) & 1);
                    mbar_wait(FULLA(sa), (it / NA) & 1);
                    mbar_wait(SPLIT(l), (it / NL) & 1);                       // A_lo written by the split warps
                    fence_after();
                    const uint64_t dah = sdesc(sbase + OFF_A + sa * A_BYTES), dal = sdesc(sbase + OFF_L + l * A_BYTES);
                    const uint64_t dbw = sdesc(sbase + OFF_W + sw * 2 * B_BYTES);     // [W1_hi chunk ; W1_lo chunk]: 128 rows
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        // columns 0..63 += A_hi W_hi, columns 64..127 += A_hi W_lo: A_hi is read from shared memory once
                        // (A itself: the tensor core ignores the low 13 mantissa bits of a TF32 operand)
                        umma_tf32(acc, dah + 2 * k, dbw + 2 * k, ID128, (kb | k) != 0);
                        umma_tf32(acc, dal + 2 * k, dbw + 2 * k, ID64, 1);            // columns 0..63 += A_lo W_hi
                    }
                    umma_commit(EMPTYA(sa));
                    umma_commit(EMPTYW(sw));
                    umma_commit(LFREE(l));
                }
                umma_commit(ACCF(b));
            }
        }
    } else if (warp < 6) {
        // ===== split warps: landed fp32 chunk -> A_lo = A - top19(A) =====
        const int t = threadIdx.x - 64;                                  // 0..127
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            for (int kb = 0; kb < NKB; ++kb, ++it) {
                const int sa = it % NA, l = it % NL;
                mbar_wait(LFREE(l), ((it / NL) & 1) ^ 1);
                mbar_wait(FULLA(sa), (it / NA) & 1);
                const uint4* A = reinterpret_cast<const uint4*>(smem + OFF_A + sa * A_BYTES);
                float4* L = reinterpret_cast<float4*>(smem + OFF_L + l * A_BYTES);
                uint4 v[A_BYTES / 16 / 128];
#pragma unroll
                for (int i = 0; i < A_BYTES / 16 / 128; ++i) v[i] = A[t + 128 * i];
#pragma unroll
                for (int i = 0; i < A_BYTES / 16 / 128; ++i) {
                    float4 lo;
                    lo.x = __uint_as_float(v[i].x) - __uint_as_float(v[i].x & 0xffffe000u);
                    lo.y = __uint_as_float(v[i].y) - __uint_as_float(v[i].y & 0xffffe000u);
                    lo.z = __uint_as_float(v[i].z) - __uint_as_float(v[i].z & 0xffffe000u);
                    lo.w = __uint_as_float(v[i].w) - __uint_as_float(v[i].w & 0xffffe000u);
                    L[t + 128 * i] = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(SPLIT(l));
            }
        }
    } else {
        // ===== epilogue warps 6..13 (TMEM lane quarter = warp % 4; two warps per quarter and per SM sub-partition, so that
        // one hides the other's shared-memory latency): layer-1 epilogue, then layers 2 and 3 in FP32 =====
        // Layers 2/3 are register-tiled like a small SGEMM (a thread per row would pull every weight through the
        // 128 B/clk shared-memory pipe for every row: 2 MB per tile): layer 2 as 8 rows x 4 columns per thread, layer 3
        // as 4 rows x 5 columns, operands from shared memory ([k][row] activations, [k][j] weights), FMAs issued as
        // packed FFMA2 (fma.rn.f32x2: same FP32 rate, half the issue slots).
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int grp = (warp - 6) >> 2;                                  // which 32 accumulator columns this warp converts
        const int et = (warp - 6) * 32 + lane;                            // 0..255
        const float* W2t = reinterpret_cast<const float*>(smem + OFF_W2T);
        const float* W3t = reinterpret_cast<const float*>(smem + OFF_W3T);
        const float* bias = reinterpret_cast<const float*>(smem + OFF_BIAS);
        float* hs = reinterpret_cast<float*>(smem + OFF_HS);
        float* outs = hs;                                                 // the result tile reuses the activation buffer
        mbar_wait(BLOB, 0);                                               // W2, W3, biases have landed (one bulk copy, under the stream)
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        const int r2 = (et >> 4) * 8, c2 = (et & 15) * 4;                 // layer-2 tile: rows r2..r2+7, columns c2..c2+3
        const int r3 = (et >> 3) * 4, c3 = (et & 7) * 5;                  // layer-3 tile: rows r3..r3+3, columns c3..c3+4
        int lt = 0;
#ifdef DART_MLP3_CLOCK
        long long ck[6] = {0, 0, 0, 0, 0, 0}, c0_ = clock64();
        const long long t_staged = c0_;
#define MLP3_CK(i) { long long t_ = clock64(); ck[i] += t_ - c0_; c0_ = t_; }
#else
#define MLP3_CK(i)
#endif
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++lt) {
            const int b = lt & 1;
            mbar_wait(ACCF(b), ((uint32_t)lt >> 1) & 1);
            fence_after();
            MLP3_CK(0)
            // layer-1 accumulator row (hi-product + lo-product halves) -> + b1 -> tanh -> hs[k][row]
#pragma unroll
            for (int cc = 0; cc < 32; cc += 16) {
                const int c0 = grp * 32 + cc;
                uint32_t r[16], rl[16];
                tmem_ld16(tl + (uint32_t)(b * 2 * HID) + c0, r);
                tmem_ld16(tl + (uint32_t)(b * 2 * HID) + HID + c0, rl);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    hs[(c0 + j) * MT + row] = tanh_fast((__uint_as_float(r[j]) + __uint_as_float(rl[j])) + bias[c0 + j]);
            }
            fence_before();
            mbar_arrive(ACCE(b));                                         // the accumulator may be overwritten (tile lt + 2)
            asm volatile("bar.sync 1, 256;" ::: "memory");               // h1 complete
            MLP3_CK(1)
            {
                unsigned long long acc[8][2];                             // [row][column pair]
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const unsigned long long bv = pack2(bias[HID + c2 + 2 * c], bias[HID + c2 + 2 * c + 1]);
#pragma unroll
                    for (int r = 0; r < 8; ++r) acc[r][c] = bv;
                }
#pragma unroll 8
                for (int k = 0; k < HID; ++k) {
                    const float4 ha = *reinterpret_cast<const float4*>(hs + k * MT + r2), hb = *reinterpret_cast<const float4*>(hs + k * MT + r2 + 4);
                    const ulonglong2 wa = *reinterpret_cast<const ulonglong2*>(W2t + k * HID + c2);
                    const float hv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const unsigned long long hh = pack2(hv[r], hv[r]);
                        ffma2(acc[r][0], hh, wa.x); ffma2(acc[r][1], hh, wa.y);
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");           // every thread has read h1: overwrite it with h2
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    float lo[8], hi[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) { unpack2(acc[r][c], lo[r], hi[r]); lo[r] = tanh_fast(lo[r]); hi[r] = tanh_fast(hi[r]); }
                    float* d0 = hs + (c2 + 2 * c) * MT + r2;
                    *reinterpret_cast<float4*>(d0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<float4*>(d0 + 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
                    *reinterpret_cast<float4*>(d0 + MT) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(d0 + MT + 4) = make_float4(hi[4], hi[5], hi[6], hi[7]);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");               // h2 complete
            MLP3_CK(2)
            {
                unsigned long long acc[2][5];                             // [row pair][column]
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const float bb = bias[2 * HID + c3 + c];
                    acc[0][c] = pack2(bb, bb); acc[1][c] = acc[0][c];
                }
#pragma unroll 8
                for (int k = 0; k < HID; ++k) {
                    const ulonglong2 hp = *reinterpret_cast<const ulonglong2*>(hs + k * MT + r3);      // rows (r3, r3+1), (r3+2, r3+3)
                    const float* wp = W3t + k * W3LD + c3;
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const unsigned long long ww = pack2(wp[c], wp[c]);
                        ffma2(acc[0][c], hp.x, ww); ffma2(acc[1][c], hp.y, ww);
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");           // every thread has read h2: overwrite it with the result tile
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const int j = c3 + c;
                    if (j < ACT) {
                        float v0, v1, v2, v3;
                        unpack2(acc[0][c], v0, v1);
                        unpack2(acc[1][c], v2, v3);
                        outs[(r3 + 0) * ACT + j] = v0; outs[(r3 + 1) * ACT + j] = v1;
                        outs[(r3 + 2) * ACT + j] = v2; outs[(r3 + 3) * ACT + j] = v3;
                    }
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            MLP3_CK(3)
            {
                const long row0 = (long)tile * MT;
                const int valid = (a.B - row0 < MT) ? (int)(a.B - row0) : MT;
                const int nflt = valid * ACT;
                float* g = a.mean + row0 * ACT;                   // 16-byte aligned: MT * ACT * 4 is a multiple of 16
                const int nv = nflt >> 2;
                for (int i = et; i < nv; i += NEPI) reinterpret_cast<float4*>(g)[i] = reinterpret_cast<const float4*>(outs)[i];
                for (int i = (nv << 2) + et; i < nflt; i += NEPI) g[i] = outs[i];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");        // the next tile's epilogue reuses hs / outs
            MLP3_CK(4)
        }
#ifdef DART_MLP3_CLOCK
        if (blockIdx.x == 0 && et == 0 && lt > 0)
            printf("mlp3 entry -> epilogue warps staged %lld cycles, -> end %lld\n", t_staged - t_entry, clock64() - t_entry);
        if (blockIdx.x == 0 && et == 0 && lt > 0)
            printf("mlp3 epilogue cycles/tile over %d tiles: wait %lld  l1-epi %lld  layer2 %lld  layer3 %lld  store %lld\n", lt, ck[0] / lt, ck[1] / lt, ck[2] / lt, ck[3] / lt, ck[4] / lt);
#endif
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(v3::TCOLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ obs build
struct ObsArgs {
    int B, count;     // count = number of pushes including this one (shared by all instances)
    const double *state, *target, *control, *cur_k;
    int ld_k;         // row stride of cur_k (34, or 36 when it aliases the LMPC aux rows)
    double *mean, *M2;
    const float* obs_in;
    float* obs_out;
};

__global__ void __launch_bounds__(256) policy_obs_kernel(const ObsArgs a) {
    constexpr int BASE = 52, HIST = 10;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)a.B * OBS) return;
    const int b = (int)(idx / OBS), j = (int)(idx % OBS);
    if (j < (HIST - 1) * BASE) {
        a.obs_out[idx] = a.obs_in[(long)b * OBS + j + BASE];
        return;
    }
    const int e = j - (HIST - 1) * BASE;
    double raw;
    if (e < 8) raw = a.state[(long)b * 8 + e];
    else if (e < 16) raw = a.target[(long)b * 8 + e - 8];
    else if (e < 18) raw = a.control[(long)b * 2 + e - 16];
    else raw = a.cur_k[(long)b * a.ld_k + e - 18];
    const double base = (double)(float)raw;                    // .astype(float32) ... .astype(float64)
    double mean = a.mean[(long)b * BASE + e], M2 = a.M2[(long)b * BASE + e];
    const double delta = base - mean;
    mean += delta / (double)a.count;
    const double delta2 = base - mean;
    M2 += delta * delta2;
    a.mean[(long)b * BASE + e] = mean;
    a.M2[(long)b * BASE + e] = M2;
    const double var = (a.count > 1) ? M2 / (double)(a.count - 1) : 1e-6;
    const float sd = (float)sqrt(fmax(var, 1e-12));
    a.obs_out[idx] = ((float)base - (float)mean) / (sd + 1e-8f);
}

// ------------------------------------------------------------------------------------------ param update
struct ParamArgs {
    int B, ld;
    const float* action;
    double* pvec;
    float k_max, max_delta, min_frac;
    double alpha, min_v, max_v;
};

__global__ void __launch_bounds__(256) policy_param_kernel(const ParamArgs a) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)a.B * ACT) return;
    const int b = (int)(idx / ACT), j = (int)(idx % ACT);
    double* slot = a.pvec + (long)b * a.ld + j;
    const double prev = *slot;
    // rlmpc2.py:745-757 in float32 (torch tensors of the action's dtype)
    const float k = (float)prev;
    const float frac = fminf(fmaxf(k / a.k_max, a.min_frac), 1.0f - 1e-6f);
    const float z_prev = logf(frac / (1.0f - frac));
    const float z_new = z_prev + a.action[idx] * a.max_delta;
    const float k_new = a.k_max * (1.0f / (1.0f + expf(-z_new)));
    // write_params_to_shm (rlmpc2.py:606-616), float64
    const double smoothed = a.alpha * (double)k_new + (1.0 - a.alpha) * prev;
    const double center = (a.max_v + a.min_v) / 2.0;
    const double scale = (a.max_v - a.min_v) / 2.0 - 1e-3;
    *slot = center + scale * tanh((smoothed - center) / scale);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeFn)p;
    return fn;
}

// 2-D fp32 row-major [rows, cols] tensor, box = [box_rows, 32 cols], 128B swizzle, zero fill out of bounds
int make_map(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeFn enc = get_encode();
    if (!enc) return DART_ERR_CUDA;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? DART_OK : DART_ERR_CUDA;
}

}  // namespace

struct dart_policy {
    int device, sms;
    float *W1, *b1, *W2, *b2, *W3, *b3;
    float *W1h, *W1l;                         // 3xTF32 split of W1 (top 19 bits / remainder)
    float* blob;                              // layers 2-3 operand block of policy_mlp3_kernel (v3::BLOB_BYTES)
    CUtensorMap tm_w1, tm_w2, tm_w3, tm_w1h, tm_w1l;
    int legacy;                               // DART_POLICY_TF32: the single-pass TF32 kernel (dart_policy_set_precision)
    int64_t launches;
};

extern "C" int dart_policy_destroy(dart_policy_handle h);

extern "C" int dart_policy_create(dart_policy_handle* out, int device, int32_t obs_dim, int32_t hidden, int32_t act_dim,
                                  const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                                  const float* b3) {
    if (!out || !W1 || !b1 || !W2 || !b2 || !W3 || !b3) return DART_ERR_ARG;
    if (obs_dim != OBS || hidden != HID || act_dim != ACT) return DART_ERR_UNSUPPORTED;   // the reference architecture
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return DART_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) return DART_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return DART_ERR_CUDA;
    dart_policy* h = new (std::nothrow) dart_policy();
    if (!h) return DART_ERR_ALLOC;
    memset(h, 0, sizeof(*h));
    h->device = device;
    cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device);
    const size_t n[6] = {(size_t)HID * OBS, HID, (size_t)HID * HID, HID, (size_t)ACT * HID, ACT};
    const size_t cap[6] = {n[0], n[1], n[2], n[3], (size_t)N3 * HID, n[5]};   // W3 zero-padded to 48 rows
    const float* src[6] = {W1, b1, W2, b2, W3, b3};
    float** dst[6] = {&h->W1, &h->b1, &h->W2, &h->b2, &h->W3, &h->b3};
    int rc = DART_OK;
    for (int i = 0; i < 6 && rc == DART_OK; ++i) {
        if (cudaMalloc(dst[i], cap[i] * sizeof(float)) != cudaSuccess) { *dst[i] = nullptr; rc = DART_ERR_ALLOC; break; }
        if (cudaMemset(*dst[i], 0, cap[i] * sizeof(float)) != cudaSuccess ||
            cudaMemcpy(*dst[i], src[i], n[i] * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
            rc = DART_ERR_CUDA;
    }
    if (rc == DART_OK) {
        // W1 = W1_hi + W1_lo: hi keeps the top 19 bits (an exact TF32 value), lo = fp32(W1 - hi), itself cut to TF32
        float* hi = new (std::nothrow) float[n[0]];
        float* lo = new (std::nothrow) float[n[0]];
        if (!hi || !lo) rc = DART_ERR_ALLOC;
        for (size_t i = 0; rc == DART_OK && i < n[0]; ++i) {
            uint32_t u;
            memcpy(&u, &W1[i], 4);
            u &= 0xffffe000u;
            memcpy(&hi[i], &u, 4);
            float l = W1[i] - hi[i];
            memcpy(&u, &l, 4);
            u &= 0xffffe000u;
            memcpy(&lo[i], &u, 4);
        }
        if (rc == DART_OK && (cudaMalloc(&h->W1h, n[0] * 4) != cudaSuccess || cudaMalloc(&h->W1l, n[0] * 4) != cudaSuccess)) rc = DART_ERR_ALLOC;
        if (rc == DART_OK && (cudaMemcpy(h->W1h, hi, n[0] * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
                              cudaMemcpy(h->W1l, lo, n[0] * 4, cudaMemcpyHostToDevice) != cudaSuccess)) rc = DART_ERR_CUDA;
        delete[] hi;
        delete[] lo;
    }
    if (rc == DART_OK) {
        // layers 2-3 operand block: W2, W3 transposed to [k][j] (W3 padded to 40 columns), then b1, b2, b3 (padded)
        float* blob = new (std::nothrow) float[v3::BLOB_BYTES / 4]();
        if (!blob) rc = DART_ERR_ALLOC;
        if (rc == DART_OK) {
            float* w2t = blob;
            float* w3t = blob + HID * HID;
            float* bs = w3t + HID * v3::W3LD;
            for (int k = 0; k < HID; ++k) {
                for (int j = 0; j < HID; ++j) w2t[k * HID + j] = W2[j * HID + k];
                for (int j = 0; j < ACT; ++j) w3t[k * v3::W3LD + j] = W3[j * HID + k];
            }
            for (int i = 0; i < HID; ++i) { bs[i] = b1[i]; bs[HID + i] = b2[i]; }
            for (int i = 0; i < ACT; ++i) bs[2 * HID + i] = b3[i];
            if (cudaMalloc(&h->blob, v3::BLOB_BYTES) != cudaSuccess) { h->blob = nullptr; rc = DART_ERR_ALLOC; }
            else if (cudaMemcpy(h->blob, blob, v3::BLOB_BYTES, cudaMemcpyHostToDevice) != cudaSuccess) rc = DART_ERR_CUDA;
        }
        delete[] blob;
    }
    if (rc == DART_OK && (make_map(&h->tm_w1, h->W1, HID, OBS, HID) || make_map(&h->tm_w2, h->W2, HID, HID, HID) ||
                          make_map(&h->tm_w3, h->W3, N3, HID, N3) || make_map(&h->tm_w1h, h->W1h, HID, OBS, HID) ||
                          make_map(&h->tm_w1l, h->W1l, HID, OBS, HID)))
        rc = DART_ERR_CUDA;
    h->legacy = 0;
    if (rc != DART_OK) { dart_policy_destroy(h); return rc; }
    if (cudaFuncSetAttribute(policy_mlp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<2>::BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(policy_mlp_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<4>::BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(policy_mlp3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v3::BYTES) != cudaSuccess) {
        dart_policy_destroy(h);
        return DART_ERR_CUDA;
    }
    *out = h;
    return DART_OK;
}

extern "C" int dart_policy_destroy(dart_policy_handle h) {
    if (!h) return DART_ERR_ARG;
    cudaSetDevice(h->device);
    float* p[9] = {h->W1, h->b1, h->W2, h->b2, h->W3, h->b3, h->W1h, h->W1l, h->blob};
    for (int i = 0; i < 9; ++i) if (p[i]) cudaFree(p[i]);
    delete h;
    return DART_OK;
}

extern "C" int dart_policy_forward(dart_policy_handle h, int32_t B, const float* obs, float* act_mean, void* stream) {
    if (!h || B < 0 || !obs || !act_mean) return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    if ((reinterpret_cast<uintptr_t>(obs) & 15) != 0 || (reinterpret_cast<uintptr_t>(act_mean) & 15) != 0)
        return DART_ERR_ARG;        // TMA rows and the coalesced output stores need 16-byte alignment
    CUtensorMap tm_obs;
    int rc = make_map(&tm_obs, obs, (uint64_t)B, OBS, MT);
    if (rc != DART_OK) return rc;
    if (!h->legacy) {
        Mlp3Args a3;
        a3.B = B; a3.ntiles = (B + MT - 1) / MT; a3.blob = h->blob;
        a3.mean = act_mean;
        const int grid = a3.ntiles < h->sms ? a3.ntiles : h->sms;           // one persistent CTA per SM
        policy_mlp3_kernel<<<grid, v3::NTHR, v3::BYTES, (cudaStream_t)stream>>>(tm_obs, h->tm_w1h, h->tm_w1l, a3);
        h->launches += 1;
        return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
    }
    MlpArgs a;
    a.B = B; a.ntiles = (B + MT - 1) / MT; a.b1 = h->b1; a.b2 = h->b2; a.b3 = h->b3; a.mean = act_mean;
    if (a.ntiles <= h->sms) {
        policy_mlp_kernel<4><<<a.ntiles, NTHREADS, Smem<4>::BYTES, (cudaStream_t)stream>>>(tm_obs, h->tm_w1, h->tm_w2, h->tm_w3, a);
    } else {
        const int grid = a.ntiles < 2 * h->sms ? a.ntiles : 2 * h->sms;     // two persistent CTAs per SM
        policy_mlp_kernel<2><<<grid, NTHREADS, Smem<2>::BYTES, (cudaStream_t)stream>>>(tm_obs, h->tm_w1, h->tm_w2, h->tm_w3, a);
    }
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" int64_t dart_policy_launch_count(dart_policy_handle h) { return h ? h->launches : -1; }

extern "C" int dart_policy_set_precision(dart_policy_handle h, int32_t precision) {
    if (!h || (precision != DART_POLICY_FP32 && precision != DART_POLICY_TF32)) return DART_ERR_ARG;
    h->legacy = precision == DART_POLICY_TF32 ? 1 : 0;
    return DART_OK;
}

extern "C" int dart_policy_obs_push(int32_t B, int32_t count, const double* state, const double* target, const double* control,
                                    const double* cur_k, int32_t ld_k, double* mean, double* M2, const float* obs_in,
                                    float* obs_out, void* stream) {
    if (B < 0 || count < 1 || !state || !target || !control || !cur_k || ld_k < ACT || !mean || !M2 || !obs_in || !obs_out ||
        obs_in == obs_out)
        return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    ObsArgs a{B, count, state, target, control, cur_k, ld_k, mean, M2, obs_in, obs_out};
    const long n = (long)B * OBS;
    policy_obs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" int dart_policy_param_update(int32_t B, const float* action, double* pvec, int32_t ld_pvec, double k_max,
                                        double max_delta, double min_k, double k_ceiling_margin, double alpha, void* stream) {
    if (B < 0 || !action || !pvec || ld_pvec < ACT || !(k_max > 0.0)) return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    ParamArgs a;
    a.B = B; a.ld = ld_pvec; a.action = action; a.pvec = pvec;
    a.k_max = (float)k_max; a.max_delta = (float)max_delta; a.min_frac = (float)(min_k / k_max);
    a.alpha = alpha; a.min_v = min_k; a.max_v = k_max - k_ceiling_margin;
    const long n = (long)B * ACT;
    policy_param_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}
