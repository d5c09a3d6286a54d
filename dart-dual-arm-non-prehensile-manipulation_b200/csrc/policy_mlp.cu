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
#include <new>

#include "../../include/dart_b200.h"

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

// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32, a=b=TF32, both K-major, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t sdesc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Production wait: try_wait suspends the thread in hardware up to a time limit, so the loop carries no extra
// instructions (round 1 kept a clock64() watchdog here: 34 % of the kernel's warp instructions were that guard).
// -DDART_MBAR_WATCHDOG restores the bounded wait (a protocol bug then traps instead of hanging) for bring-up.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#ifdef DART_MBAR_WATCHDOG
    const long long t0 = clock64();
    uint32_t done = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
#else
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(bar), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

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
    CUtensorMap tm_w1, tm_w2, tm_w3;
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
    if (rc == DART_OK && (make_map(&h->tm_w1, h->W1, HID, OBS, HID) || make_map(&h->tm_w2, h->W2, HID, HID, HID) ||
                          make_map(&h->tm_w3, h->W3, N3, HID, N3)))
        rc = DART_ERR_CUDA;
    if (rc != DART_OK) { dart_policy_destroy(h); return rc; }
    if (cudaFuncSetAttribute(policy_mlp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<2>::BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(policy_mlp_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<4>::BYTES) != cudaSuccess) {
        dart_policy_destroy(h);
        return DART_ERR_CUDA;
    }
    *out = h;
    return DART_OK;
}

extern "C" int dart_policy_destroy(dart_policy_handle h) {
    if (!h) return DART_ERR_ARG;
    cudaSetDevice(h->device);
    float* p[6] = {h->W1, h->b1, h->W2, h->b2, h->W3, h->b3};
    for (int i = 0; i < 6; ++i) if (p[i]) cudaFree(p[i]);
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
