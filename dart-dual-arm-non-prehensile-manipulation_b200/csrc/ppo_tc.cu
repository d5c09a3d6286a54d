// The two layer-1 GEMMs of the PPO step on the 5th-generation tensor cores (tcgen05.mma kind::tf32, FP32 accumulators in
// TMEM) -- 81 % of the training step's flops (LMPC/src/controller/rlmpc2.py:801 forward, :815 backward of Policy.mean_net[0] /
// value_net[0]; both networks at once: the combined W1 is [128, 520]):
//
//   ppo_l1_fwd_tc_kernel    Z1[M,128]   = X[M,520] W1^T            (pre-activation; bias + tanh stay in ppo_mid_kernel)
//   ppo_l1_wgrad_tc_kernel  dW1[128,520] = dZ1^T[128,M] X[M,520]   (+ db1 = column sums of dZ1), split over the minibatch
//
// FP32 fidelity (the reference trains in FP32, torch never enables TF32): every product is the 3xTF32 sum
//     a b = a_hi b_hi + a_hi b_lo + a_lo b_hi,   hi = top 19 bits, lo = fp32(x - hi)   (dropped term ~ 2^-22 relative),
// with FP32 accumulation in TMEM, so gradients stay inside the bounds of tests/test_gpu_ppo.py (2e-4 of the tensor's max).
//
// No TMA here: the minibatch is a PERMUTATION of the pooled rollout (rows gathered by index), W1 changes every optimiser step,
// and the weight gradient contracts over the SAMPLE index, i.e. needs both operands transposed.  Eight loader warps therefore
// build the operand tiles themselves: 16-byte global loads (gathered rows), hi/lo split in registers, stores straight into
// the K-major 128-byte-swizzled layout the UMMA descriptors expect (tc_common.cuh), then a generic->async proxy fence and an
// mbarrier arrive; one thread issues the MMAs; tcgen05.commit hands each stage back to the loaders.  Both kernels are one
// CTA per SM (192-200 kB of operand stages), 128 accumulator rows = the 128 TMEM lanes.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dart_b200.h"
#include "ppo_tc.h"
#include "tc_common.cuh"

namespace {
using namespace dart_tc;

constexpr int OBS = 520, H2W = 128;
constexpr int BK = 32;                         // fp32 per 128-byte swizzle row = K extent of one stage
constexpr int ROW_B = 128;                     // bytes per operand-tile row
constexpr int NLOAD = 256;                     // loader threads (warps 1..8); warp 0 issues the MMAs
constexpr int NTHR = 32 + NLOAD;

__device__ __forceinline__ void split3(const float4 v, uint4& h, float4& l) {
    h.x = __float_as_uint(v.x) & 0xffffe000u; h.y = __float_as_uint(v.y) & 0xffffe000u;
    h.z = __float_as_uint(v.z) & 0xffffe000u; h.w = __float_as_uint(v.w) & 0xffffe000u;
    l.x = v.x - __uint_as_float(h.x); l.y = v.y - __uint_as_float(h.y);
    l.z = v.z - __uint_as_float(h.z); l.w = v.w - __uint_as_float(h.w);
}
// byte offset of the 16-byte chunk `c` (0..7) of tile row `r`
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)(r * ROW_B + ((c ^ (r & 7)) << 4)); }

// ============================================================================================== forward
// Measured on B200: the FP32 accumulation of tcgen05.mma in TMEM TRUNCATES (every accumulate step loses up to one ulp of the
// running sum, always towards zero), so a K = 520 contraction accumulated entirely in TMEM carries a systematic ~3e-6 relative
// bias -- harmless for the policy's evaluation forward, but the PPO ratio exp(logp - logp_old) with std = 0.1 amplifies a
// forward error ~100x into the actor's gradient (1.7e-3 against autograd, bound 1e-4).  The forward therefore PROMOTES: the
// dominant product A_hi W_hi of each 32-wide K chunk goes to one of two TMEM buffers (4 accumulate steps), four promotion
// warps add the finished chunk into FP32 registers with round-to-nearest (a thread per row, 128 columns) while the next chunk
// is multiplied; the two correction products (2^-11 of the magnitude, their truncation is below FP32 noise) accumulate in
// TMEM over the whole K range and are added once at the end.
namespace fwd {
constexpr int MT = 128;                                         // rows per CTA = UMMA M
constexpr int NKB = (OBS + BK - 1) / BK;                        // 17 K chunks, the last one zero-filled past column 520
constexpr int TILE = MT * ROW_B;                                // 16 kB: A_hi | A_lo | W_hi | W_lo
constexpr int STAGE = 4 * TILE;
constexpr int S = 3;
constexpr int OFF_BAR = S * STAGE;
constexpr int NBAR = 2 * S + 5;                                 // full[S] empty[S] accf[2] acce[2] corr
constexpr int BYTES = OFF_BAR + 8 * NBAR + 16 + 1024;
constexpr uint32_t TCOLS = 512;                                 // [0,128) [128,256): chunk buffers of A_hi W_hi; [256,384): corrections
constexpr uint32_t CORR = 256;
constexpr int NPROM = 256;                                      // promotion threads: warps 9..16, two per TMEM lane quarter (64 columns each)
constexpr int NTHR_F = 32 + NLOAD + NPROM;
}  // namespace fwd

struct FwdArgs {
    int M;
    const float* X;          // [rows, 520]
    const int64_t* gidx;     // optional row gather
    const float* W1;         // [128, 520]
    float* Z;                // [M, 128]
};

__global__ void __launch_bounds__(fwd::NTHR_F, 1) ppo_l1_fwd_tc_kernel(const FwdArgs a) {
    using namespace fwd;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = sbase + OFF_BAR;
    auto FULL = [&](int s) { return bar0 + 8 * s; };
    auto EMPTY = [&](int s) { return bar0 + 8 * (S + s); };
    auto ACCF = [&](int b) { return bar0 + 8 * (2 * S + b); };          // chunk buffer b holds a finished chunk
    auto ACCE = [&](int b) { return bar0 + 8 * (2 * S + 2 + b); };      // chunk buffer b has been added into the registers
    const uint32_t CORRF = bar0 + 8 * (2 * S + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * NBAR);
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(FULL(s), NLOAD); mbar_init(EMPTY(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(ACCF(b), 1); mbar_init(ACCE(b), NPROM); }
        mbar_init(CORRF, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const int row0 = blockIdx.x * MT;

    if (warp == 0) {
        if (lane == 0) {
            constexpr uint32_t ID128 = idesc_tf32(MT, H2W);
            for (int kb = 0; kb < NKB; ++kb) {
                const int s = kb % S, b = kb & 1;
                mbar_wait(FULL(s), (uint32_t)(kb / S) & 1);
                mbar_wait(ACCE(b), (((uint32_t)kb >> 1) & 1) ^ 1);        // chunk kb - 2 has left this buffer
                fence_after();
                const uint32_t st = sbase + s * STAGE;
                const uint64_t dah = sdesc(st), dal = sdesc(st + TILE), dwh = sdesc(st + 2 * TILE), dwl = sdesc(st + 3 * TILE);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    umma_tf32(tmem + (uint32_t)(b * H2W), dah + 2 * k, dwh + 2 * k, ID128, k != 0);      // this chunk's A_hi W_hi
                    umma_tf32(tmem + CORR, dah + 2 * k, dwl + 2 * k, ID128, (kb | k) != 0);               // corrections, all chunks
                    umma_tf32(tmem + CORR, dal + 2 * k, dwh + 2 * k, ID128, 1);
                }
                umma_commit(EMPTY(s));
                umma_commit(ACCF(b));
            }
            umma_commit(CORRF);
        }
    } else if (warp <= 8) {
        // ===== loaders: thread t owns chunk column (t & 7) of tile rows (t >> 3) + 32 i, for the X tile and the W1 tile =====
        const int t = threadIdx.x - 32;
        const int c = t & 7;
        const float* xrow[4];
        const float* wrow[4];
        uint32_t off[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = (t >> 3) + 32 * i;
            const long gr = (long)row0 + r;
            xrow[i] = gr < a.M ? a.X + (size_t)(a.gidx ? a.gidx[gr] : gr) * OBS : nullptr;
            wrow[i] = a.W1 + (size_t)r * OBS;
            off[i] = sw128(r, c);
        }
        auto fetch = [&](int kb, float4 (&xv)[4], float4 (&wv)[4]) {
            const int col = kb * BK + 4 * c;
            const bool cv = col < OBS;                                   // 520 = 16 * 32 + 8: whole 16-byte chunks are in or out
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                xv[i] = (cv && xrow[i]) ? __ldg(reinterpret_cast<const float4*>(xrow[i] + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
                wv[i] = cv ? __ldg(reinterpret_cast<const float4*>(wrow[i] + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        float4 xv[4], wv[4], xn[4], wn[4];
        fetch(0, xv, wv);
        for (int kb = 0; kb < NKB; ++kb) {
            const int s = kb % S;
            if (kb + 1 < NKB) fetch(kb + 1, xn, wn);                     // the next chunk's loads fly under this chunk's stores
            mbar_wait(EMPTY(s), ((uint32_t)(kb / S) & 1) ^ 1);
            uint8_t* st = smem + s * STAGE;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint4 h; float4 l;
                split3(xv[i], h, l);
                *reinterpret_cast<uint4*>(st + off[i]) = h;
                *reinterpret_cast<float4*>(st + TILE + off[i]) = l;
                split3(wv[i], h, l);
                *reinterpret_cast<uint4*>(st + 2 * TILE + off[i]) = h;
                *reinterpret_cast<float4*>(st + 3 * TILE + off[i]) = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(FULL(s));
#pragma unroll
            for (int i = 0; i < 4; ++i) { xv[i] = xn[i]; wv[i] = wn[i]; }
        }
    } else {
        // ===== promotion warps 9..16 (TMEM lane quarter = warp % 4, column half = (warp - 9) / 4): half a Z row in FP32
        // registers, round-to-nearest adds =====
        constexpr int HC = H2W / 2;
        const int q = warp & 3, half = (warp - 9) >> 2;
        const int r = q * 32 + lane;
        const long gr = (long)row0 + r;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HC);
        float acc[HC];
#pragma unroll
        for (int j = 0; j < HC; ++j) acc[j] = 0.f;
        for (int kb = 0; kb < NKB; ++kb) {
            const int b = kb & 1;
            mbar_wait(ACCF(b), ((uint32_t)kb >> 1) & 1);
            fence_after();
#pragma unroll
            for (int c0 = 0; c0 < HC; c0 += 16) {
                uint32_t u[16];
                tmem_ld16(tl + (uint32_t)(b * H2W) + c0, u);
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(u[j]);
            }
            fence_before();
            mbar_arrive(ACCE(b));
        }
        mbar_wait(CORRF, 0);
        fence_after();
#pragma unroll
        for (int c0 = 0; c0 < HC; c0 += 16) {
            uint32_t u[16];
            tmem_ld16(tl + CORR + c0, u);
            if (gr < a.M) {
                float4* dst = reinterpret_cast<float4*>(a.Z + (size_t)gr * H2W + half * HC + c0);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    dst[j] = make_float4(acc[c0 + 4 * j] + __uint_as_float(u[4 * j]), acc[c0 + 4 * j + 1] + __uint_as_float(u[4 * j + 1]),
                                         acc[c0 + 4 * j + 2] + __uint_as_float(u[4 * j + 2]), acc[c0 + 4 * j + 3] + __uint_as_float(u[4 * j + 3]));
            }
        }
        fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(fwd::TCOLS) : "memory");
    }
}

// ============================================================================================== weight gradient
namespace wg {
constexpr int NS = 144;                                         // feature columns per CTA (slab) = UMMA N: 520 -> 3 x 144 + 88 (+ 56 pad rows)
constexpr int NSLAB = (OBS + NS - 1) / NS;                      // 4
constexpr int A_T = H2W * ROW_B;                                // 16 kB: dZ1^T tile, 128 neurons x 32 samples
constexpr int B_T = NS * ROW_B;                                 // 18 kB: X^T tile, 144 features x 32 samples
constexpr int STAGE = 2 * A_T + 2 * B_T;                        // A_hi | A_lo | B_hi | B_lo = 68 kB
constexpr int S = 3;
constexpr int OFF_BAR = S * STAGE;
constexpr int BYTES = OFF_BAR + 8 * (2 * S + 1) + 16 + 1024;
constexpr uint32_t TCOLS = 256;                                 // 144 accumulator columns (allocation is a power of two)
constexpr int ONES_ROW = OBS - (NSLAB - 1) * NS;                // last slab, tile row 88: a row of ones -> column sums of dZ1 (db1)
constexpr int APAIRS = H2W / 8, BPAIRS = NS / 8;                // loader work units: 8 consecutive columns of one sample (32 bytes)
constexpr int UNITS = APAIRS + BPAIRS, UPW = (UNITS + 7) / 8;   // 34 units over 8 loader warps
static_assert(ONES_ROW % 8 == 0 && ONES_ROW < NS, "the ones row is the first pad row of the last slab");
}  // namespace wg

struct WgArgs {
    int M, per;              // samples, samples per split (multiple of 32)
    const float* dZ;         // [M, 128]
    const float* X;          // [rows, 520]
    const int64_t* gidx;
    float* part;             // split s writes dW1 at part + s * stride (row-major [128, 520]) and db1 at part + s * stride + 128 * 520
    long stride;
};

// The TMEM accumulation truncates (see the forward kernel); here a split accumulates at most `per` / 8 * 3 steps (a few
// hundred) of products whose sum is not amplified downstream: measured 2.7e-5 against autograd, the FP32 SIMT kernel 1.4e-5,
// torch's own FP32 result 2.0e-5 against float64.  The splits are summed in FP32 with round-to-nearest by the reduce kernel.
__global__ void __launch_bounds__(NTHR, 1) ppo_l1_wgrad_tc_kernel(const WgArgs a) {
    using namespace wg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = sbase + OFF_BAR;
    auto FULL = [&](int s) { return bar0 + 8 * s; };
    auto EMPTY = [&](int s) { return bar0 + 8 * (S + s); };
    const uint32_t ACCF = bar0 + 8 * (2 * S);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * (2 * S + 1));
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(FULL(s), NLOAD); mbar_init(EMPTY(s), 1); }
        mbar_init(ACCF, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const int slab = blockIdx.x, split = blockIdx.y;
    const int f0 = slab * NS;
    const int width = (OBS - f0 < NS) ? OBS - f0 : NS;              // valid feature columns of this slab (144 or 88)
    const bool ones = slab == NSLAB - 1;
    const int s_begin = split * a.per;
    const int s_end = (s_begin + a.per < a.M) ? s_begin + a.per : a.M;
    const int nchunk = (s_end - s_begin + BK - 1) / BK;             // >= 1 (the host sizes the grid so)

    if (warp == 0) {
        if (lane == 0) {
            constexpr uint32_t ID = idesc_tf32(H2W, NS);
            for (int ch = 0; ch < nchunk; ++ch) {
                const int s = ch % S;
                mbar_wait(FULL(s), (uint32_t)(ch / S) & 1);
                fence_after();
                const uint32_t st = sbase + s * STAGE;
                const uint64_t dah = sdesc(st), dal = sdesc(st + A_T), dbh = sdesc(st + 2 * A_T), dbl = sdesc(st + 2 * A_T + B_T);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    umma_tf32(tmem, dah + 2 * k, dbh + 2 * k, ID, (ch | k) != 0);
                    umma_tf32(tmem, dah + 2 * k, dbl + 2 * k, ID, 1);
                    umma_tf32(tmem, dal + 2 * k, dbh + 2 * k, ID, 1);
                }
                umma_commit(EMPTY(s));
            }
            umma_commit(ACCF);
        }
    } else {
        // ===== loaders: lane = sample of the chunk (the K index of both operands), warps split the 8-column units; each unit is
        // two 16-byte loads of one 32-byte sector, split hi/lo, and 2 x 8 scalar stores: the 32 lanes of a warp write one tile
        // row (32 consecutive K entries, permuted in 16-byte chunks by the swizzle) -- conflict-free.  All loads of a chunk are
        // issued before the first store (the loop was latency-bound with one unit in flight) =====
        const int lw = warp - 1;                                       // 0..7
        const int kc = lane >> 2, kw = (lane & 3) * 4;                 // this sample's 16-byte chunk and byte offset inside a tile row
        // one chunk's 16-byte loads (gathered row pointer first) into registers
        auto fetch = [&](int ch, float4 (&v)[UPW][2]) {
            const int smp = s_begin + ch * BK + lane;
            const bool sv = smp < s_end;
            const float* dz = a.dZ + (size_t)(sv ? smp : 0) * H2W;
            const float* xr = a.X + (size_t)(sv ? (a.gidx ? a.gidx[smp] : smp) : 0) * OBS + f0;
#pragma unroll
            for (int j = 0; j < UPW; ++j) {
                const int u = lw + 8 * j;
                const bool isA = u < APAIRS;
                const int r0 = isA ? u * 8 : (u - APAIRS) * 8;           // first tile row (neuron / feature) of the unit
                const float* src = isA ? dz + r0 : xr + r0;
                v[j][0] = make_float4(0.f, 0.f, 0.f, 0.f); v[j][1] = v[j][0];
                if (u < UNITS && sv && (isA || r0 < width)) {            // width is a multiple of 8: units are in or out
                    v[j][0] = __ldg(reinterpret_cast<const float4*>(src));
                    v[j][1] = __ldg(reinterpret_cast<const float4*>(src + 4));
                }
                if (!isA && ones && r0 == ONES_ROW && sv) v[j][0].x = 1.0f;      // the ones row (first pad row of the last slab)
            }
        };
        float4 v[UPW][2], vn[UPW][2];
        fetch(0, v);
        for (int ch = 0; ch < nchunk; ++ch) {
            const int s = ch % S;
            if (ch + 1 < nchunk) fetch(ch + 1, vn);                      // the next chunk's loads fly under this chunk's stores
            mbar_wait(EMPTY(s), ((uint32_t)(ch / S) & 1) ^ 1);
            uint8_t* st = smem + s * STAGE;
#pragma unroll
            for (int j = 0; j < UPW; ++j) {
                const int u = lw + 8 * j;
                if (u >= UNITS) continue;
                const bool isA = u < APAIRS;
                const int r0 = isA ? u * 8 : (u - APAIRS) * 8;
                uint4 h0, h1; float4 l0, l1;
                split3(v[j][0], h0, l0);
                split3(v[j][1], h1, l1);
                uint8_t* hi = st + (isA ? 0 : 2 * A_T);
                uint8_t* lo = st + (isA ? A_T : 2 * A_T + B_T);
                const uint32_t hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                const float lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t o = sw128(r0 + i, kc) + kw;
                    *reinterpret_cast<uint32_t*>(hi + o) = hv[i];
                    *reinterpret_cast<float*>(lo + o) = lv[i];
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(FULL(s));
#pragma unroll
            for (int j = 0; j < UPW; ++j) { v[j][0] = vn[j][0]; v[j][1] = vn[j][1]; }
        }
        // ===== epilogue (warps 1..4): this split's partial dW1 rows (neuron = TMEM lane) and db1 =====
        if (warp <= 4) {
            mbar_wait(ACCF, 0);
            fence_after();
            const int q = warp & 3;
            const int n = q * 32 + lane;
            const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
            float* out = a.part + (size_t)split * a.stride + (size_t)n * OBS + f0;
#pragma unroll 1
            for (int c0 = 0; c0 < NS; c0 += 16) {
                if (c0 >= width && !(ones && c0 <= ONES_ROW)) break;
                uint32_t u[16];
                tmem_ld16(tl + c0, u);
                if (c0 + 16 <= width) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        reinterpret_cast<float4*>(out + c0)[j] = make_float4(__uint_as_float(u[4 * j]), __uint_as_float(u[4 * j + 1]),
                                                                             __uint_as_float(u[4 * j + 2]), __uint_as_float(u[4 * j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (c0 + j < width) out[c0 + j] = __uint_as_float(u[j]);
                        else if (ones && c0 + j == ONES_ROW) a.part[(size_t)split * a.stride + (size_t)H2W * OBS + n] = __uint_as_float(u[j]);
                    }
                }
            }
            fence_before();
        }
    }
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(wg::TCOLS) : "memory");
    }
}

}  // namespace

namespace dart_ppo_tc {

static int ensure_attrs() {
    static int set_dev[64] = {0};      // the opt-in shared-memory size is per device function and per device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return DART_ERR_CUDA;
    if (!set_dev[dev]) {
        if (cudaFuncSetAttribute(ppo_l1_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd::BYTES) != cudaSuccess) return DART_ERR_CUDA;
        if (cudaFuncSetAttribute(ppo_l1_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::BYTES) != cudaSuccess) return DART_ERR_CUDA;
        set_dev[dev] = 1;
    }
    return DART_OK;
}

int l1_forward(int M, const float* X, const int64_t* gidx, const float* W1, float* Z, cudaStream_t st) {
    if (ensure_attrs() != DART_OK) return DART_ERR_CUDA;
    FwdArgs a{M, X, gidx, W1, Z};
    ppo_l1_fwd_tc_kernel<<<(M + fwd::MT - 1) / fwd::MT, fwd::NTHR_F, fwd::BYTES, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

int l1_wgrad_splits(int M, int max_splits, int* per_out) {
    // whole 32-sample chunks per split; every split non-empty
    int splits = max_splits < 1 ? 1 : max_splits;
    int per = (M + splits - 1) / splits;
    per = (per + BK - 1) / BK * BK;
    splits = (M + per - 1) / per;
    *per_out = per;
    return splits;
}

int l1_wgrad(int M, int splits, int per, const float* dZ, const float* X, const int64_t* gidx, float* part, long stride, cudaStream_t st) {
    if (ensure_attrs() != DART_OK) return DART_ERR_CUDA;
    WgArgs a{M, per, dZ, X, gidx, part, stride};
    ppo_l1_wgrad_tc_kernel<<<dim3(wg::NSLAB, splits), NTHR, wg::BYTES, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

}  // namespace dart_ppo_tc
