// PPO training path of the LMPC parameter-adaptation policy, batched over instances.
//
// Replaces, for B instances at once, what RLMPC._rl_worker does with torch autograd for one instance
// (LMPC/src/controller/rlmpc2.py):
//   ppo_sample_kernel     rollout-time action:  a = mean + std * eps, log-probability, value        (:670-699)
//   ppo_reward_kernel     proximity reward, penalties, bonuses, termination                            (:598-601, 701-735)
//   ppo_gae_kernel        generalised advantage estimation along each instance's rollout, returns     (:589-596, 783-784)
//   ppo_normalize_kernel  (x - mean) / (std + 1e-8) over the pooled rollout (numpy std for returns :785, torch std :792)
//   ppo_gather_kernel     minibatch gather by a permutation                                            (:795-800)
//   gemm_kernel           FP32 SIMT tiled GEMM, grouped over the actor/critic pair: forward (bias + tanh), data
//                         gradient (fused tanh'), weight gradient (split over the minibatch, partials summed in a
//                         fixed order => bitwise repeatable)                                           (:801, :815)
//   ppo_loss_kernel       clipped surrogate + value MSE + entropy bonus and their gradients w.r.t. mean/value/log_std (:802-813)
//   ppo_grad_reduce_kernel, ppo_adam_kernel   clip_grad_norm_(0.5) and Adam with L2 weight decay       (:561, :814-817)
//
// Arithmetic is FP32 like the reference's torch modules (per-sample scalars and all reductions in FP64).  This is the
// first correct device path of the training step: the GEMMs run on the FP32 CUDA-core pipe, not on tcgen05, so that the
// gradients match torch's FP32 autograd (TF32 would not).  The reference's rollout buffer swaps reward and value
// (`buf.add(..., value, reward, ...)` against `add(o, a, logp, r, v, done)`, :744 vs :93); that bug is not reproduced.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <new>

#include "../../include/dart_b200.h"
#include "ppo_tc.h"

namespace {

constexpr int OBS = 520, HID = 64, ACT = 34, H2W = 2 * HID;
// flat parameter vector (float32): every weight matrix is followed by its bias, torch layout [out, in]
constexpr int OFF_W1 = 0;                          // [128,520] rows 0-63 mean_net.0.weight, rows 64-127 value_net.0.weight
constexpr int OFF_B1 = OFF_W1 + H2W * OBS;         // [128]
constexpr int OFF_W2A = OFF_B1 + H2W;              // mean_net.2.weight [64,64]
constexpr int OFF_B2A = OFF_W2A + HID * HID;
constexpr int OFF_W2C = OFF_B2A + HID;             // value_net.2.weight [64,64]
constexpr int OFF_B2C = OFF_W2C + HID * HID;
constexpr int OFF_W3A = OFF_B2C + HID;             // mean_net.4.weight [34,64]
constexpr int OFF_B3A = OFF_W3A + ACT * HID;
constexpr int OFF_W3C = OFF_B3A + ACT;             // value_net.4.weight [1,64]
constexpr int OFF_B3C = OFF_W3C + HID;
constexpr int OFF_LS = OFF_B3C + 1;                // log_std [34]
constexpr int NPW = OFF_LS;                        // weights + biases
constexpr int NP = OFF_LS + ACT;                   // 77 317 parameters

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4, GT = 256;
constexpr int LOSS_T = 128, LOSS_W = 3 + ACT;      // per-block partials: policy sum, value sq. sum, unused, dlog_std[34]
constexpr int MAX_SPLITS = 64;                     // layer-1 weight gradient: chunks of >= 256 samples
constexpr int MAX_SPLITS_SMALL = 256;              // layers 2-3: chunks of >= 64 samples (more CTAs in flight for these latency-bound GEMMs)
constexpr int NPB = OFF_W2A;                       // layer-1 block of the flat vector (W1 | b1)
constexpr int NPS = NPW - OFF_W2A;                 // layers 2-3 block
constexpr int NPS_LD = (NPS + 3) / 4 * 4;          // row stride of its partial buffer (16-byte aligned rows)
constexpr int FWD_KCHUNK = 80, FWD_SPLITS = 7;     // layer-1 forward of small minibatches: 520 = 6 x 80 + 40
constexpr int BIG_FWD_SPLITS = 2, BIG_FWD_KCHUNK = 264;   // layer-1 forward of large minibatches: two K halves (multiple of 8)
constexpr int BIG_MIN_ROWS = 512;                  // minibatches from here on use the 128x128 kernel for the layer-1 GEMMs
constexpr double HALF_LOG_2PI = 0.91893853320467274178;

// One GEMM of a grouped launch:  C[m,n] = epilogue( sum_k A(m,k) B(k,n) ).
struct GemmProb {
    const float* A; int lda; int ta;      // A(m,k) = ta ? A[k*lda + m] : A[m*lda + k]
    const float* B; int ldb; int tb;      // B(k,n) = tb ? B[n*ldb + k] : B[k*ldb + n]
    float* C; int ldc;
    int M, N, K;
    int mode;                             // 0 store, 1 tanh(acc + bias[n]), 2 acc + bias[n], 3 acc * (1 - H[m,n]^2)
    const float* bias;
    const float* H; int ldh;
    float* rowsum;                        // split mode: partial sums over k of A(m,k) (the bias gradient), or nullptr
    const int64_t* gidx; int gmode;       // minibatch gather without a copy: 1 = A's row m -> gidx[m] (ta = 0), 2 = B's k -> gidx[k] (tb = 0)
};
struct GemmGroup {
    GemmProb p[2];
    int count, splits, kchunk;
    long split_stride;                    // floats between the outputs of consecutive splits
};

__global__ void __launch_bounds__(GT) gemm_kernel(const GemmGroup g) {
    const int pi = blockIdx.z / g.splits, split = blockIdx.z % g.splits;
    const GemmProb& P = g.p[pi];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= P.M || n0 >= P.N) return;
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    const int t = threadIdx.x, ty = t / 16, tx = t % 16;
    const int k_begin = split * g.kchunk;
    const int k_end = min(P.K, k_begin + g.kchunk);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    const bool want_rs = P.rowsum != nullptr && blockIdx.x == 0 && tx == 0;
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
        for (int i = 0; i < (BM * BK) / GT; ++i) {
            const int e = t + i * GT;
            int m, k;
            if (P.ta) { k = e / BM; m = e % BM; } else { m = e / BK; k = e % BK; }
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < P.M && gk < k_end) v = P.ta ? P.A[(size_t)gk * P.lda + gm] : P.A[(size_t)(P.gmode == 1 ? P.gidx[gm] : gm) * P.lda + gk];
            As[k][m] = v;
        }
#pragma unroll
        for (int i = 0; i < (BN * BK) / GT; ++i) {
            const int e = t + i * GT;
            int n, k;
            if (P.tb) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < P.N && gk < k_end) v = P.tb ? P.B[(size_t)gn * P.ldb + gk] : P.B[(size_t)(P.gmode == 2 ? P.gidx[gk] : gk) * P.ldb + gn];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            if (want_rs) {
#pragma unroll
                for (int i = 0; i < 4; ++i) rs[i] += av[i];
            }
        }
        __syncthreads();
    }
    float* C = P.C + (size_t)split * g.split_stride;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= P.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= P.N) continue;
            float v = acc[i][j];
            if (P.mode == 1) v = tanhf(v + P.bias[n]);
            else if (P.mode == 2) v = v + P.bias[n];
            else if (P.mode == 3) { const float h = P.H[(size_t)m * P.ldh + n]; v = v * (1.f - h * h); }
            C[(size_t)m * P.ldc + n] = v;
        }
        if (want_rs) P.rowsum[(size_t)split * g.split_stride + m] = rs[i];
    }
}

// The two layer-1 GEMMs (forward [M,520]x[520,128] and weight gradient [128,M]x[M,520]) carry 81 % of the step's flops:
// 128x128 tiles, BK = 8, 8x8 register micro-tile per thread (as 2x2 blocks of 4x4 so that the shared-memory reads are
// conflict-free float4s), 16-byte global loads, next tile prefetched into registers while the current one is multiplied
// (two shared-memory buffers, one barrier per k-step).  Same operand conventions, epilogues and split/rowsum outputs
// as gemm_kernel; requires 16-byte aligned operands with leading dimensions that are multiples of 4.
constexpr int TM = 128, TN = 128, TK = 8, TP = 4;
__device__ __forceinline__ void load_tile(const float* __restrict__ X, int ld, bool trans_rows_contig, int r0, int R, int k0,
                                          int k_end, int t, float (&v)[4], int& r, int& k, const int64_t* gidx = nullptr) {
    // gidx: gather on the sample index (the row r when k is contiguous, the column k when r is contiguous)
    // trans_rows_contig: X(r,k) = X[k*ld + r] (r contiguous) else X[r*ld + k] (k contiguous); returns 4 elements:
    // rows r..r+3 at column k (contiguous r) or row r at columns k..k+3 (contiguous k)
    if (trans_rows_contig) { k = t / 32; r = (t % 32) * 4; } else { r = t / 2; k = (t % 2) * 4; }
    const int gr = r0 + r, gk = k0 + k;
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (trans_rows_contig) {
        if (gk < k_end) {
            const float* src = X + (size_t)(gidx ? gidx[gk] : gk) * ld + gr;
            if (gr + 3 < R) { const float4 q = *reinterpret_cast<const float4*>(src); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
            else { for (int i = 0; i < 4; ++i) if (gr + i < R) v[i] = src[i]; }
        }
    } else {
        if (gr < R) {
            const float* src = X + (size_t)(gidx ? gidx[gr] : gr) * ld + gk;
            if (gk + 3 < k_end) { const float4 q = *reinterpret_cast<const float4*>(src); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
            else { for (int i = 0; i < 4; ++i) if (gk + i < k_end) v[i] = src[i]; }
        }
    }
}
__device__ __forceinline__ void store_tile(float (*S)[TM + TP], bool rows_contig, int r, int k, const float (&v)[4]) {
    if (rows_contig) *reinterpret_cast<float4*>(&S[k][r]) = make_float4(v[0], v[1], v[2], v[3]);
    else { S[k][r] = v[0]; S[k + 1][r] = v[1]; S[k + 2][r] = v[2]; S[k + 3][r] = v[3]; }
}

__global__ void __launch_bounds__(GT, 2) gemm128_kernel(const GemmProb P, const int splits, const int kchunk, const long split_stride) {
    const int split = blockIdx.z;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    __shared__ __align__(16) float As[2][TK][TM + TP];
    __shared__ __align__(16) float Bs[2][TK][TN + TP];
    const int t = threadIdx.x, ty = t / 16, tx = t % 16;
    const int k_begin = split * kchunk;
    const int k_end = min(P.K, k_begin + kchunk);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float rs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool want_rs = P.rowsum != nullptr && blockIdx.x == 0 && tx == 0;
    const bool a_rc = P.ta != 0, b_rc = P.tb == 0;      // "rows contiguous": A(m,k) with m contiguous / B(k,n) with n contiguous
    const int64_t* ga = P.gmode == 1 ? P.gidx : nullptr;
    const int64_t* gb = P.gmode == 2 ? P.gidx : nullptr;
    float va[4], vb[4];
    int ar, ak, br, bk;
    int buf = 0;
    if (k_begin < k_end) {
        load_tile(P.A, P.lda, a_rc, m0, P.M, k_begin, k_end, t, va, ar, ak, ga);
        load_tile(P.B, P.ldb, b_rc, n0, P.N, k_begin, k_end, t, vb, br, bk, gb);
        store_tile(As[0], a_rc, ar, ak, va);
        store_tile(Bs[0], b_rc, br, bk, vb);
    }
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += TK) {
        const bool more = k0 + TK < k_end;
        if (more) {
            load_tile(P.A, P.lda, a_rc, m0, P.M, k0 + TK, k_end, t, va, ar, ak, ga);
            load_tile(P.B, P.ldb, b_rc, n0, P.N, k0 + TK, k_end, t, vb, br, bk, gb);
        }
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            if (want_rs) {
#pragma unroll
                for (int i = 0; i < 8; ++i) rs[i] += av[i];
            }
        }
        if (more) {
            store_tile(As[buf ^ 1], a_rc, ar, ak, va);
            store_tile(Bs[buf ^ 1], b_rc, br, bk, vb);
        }
        __syncthreads();
        buf ^= 1;
    }
    float* C = P.C + (size_t)split * split_stride;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= P.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= P.N) continue;
            float v = acc[i][j];
            if (P.mode == 1) v = tanhf(v + P.bias[n]);
            else if (P.mode == 2) v = v + P.bias[n];
            else if (P.mode == 3) { const float h = P.H[(size_t)m * P.ldh + n]; v = v * (1.f - h * h); }
            C[(size_t)m * P.ldc + n] = v;
        }
        if (want_rs) P.rowsum[(size_t)split * split_stride + m] = rs[i];
    }
}

// ---- rollout-time sampling (rlmpc2.py:670-699): a = mean + std * eps, logp = sum_j log N(a_j; mean_j, std_j) ----
__global__ void ppo_sample_kernel(int B, const float* __restrict__ mean, const float* __restrict__ log_std, float ls_min,
                                  float ls_max, const float* __restrict__ eps, float* __restrict__ action,
                                  float* __restrict__ logp) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double lp = 0.0;
    for (int j = 0; j < ACT; ++j) {
        const float ls = fminf(fmaxf(log_std[j], ls_min), ls_max);
        const float sd = fmaxf(expf(ls), 1e-6f);
        const float e = eps ? eps[(size_t)b * ACT + j] : 0.f;
        const float mu = mean[(size_t)b * ACT + j];
        const float a = fmaf(sd, e, mu);
        action[(size_t)b * ACT + j] = a;
        const double z = ((double)a - (double)mu) / (double)sd;
        lp += -0.5 * z * z - (double)logf(sd) - HALF_LOG_2PI;
    }
    logp[b] = (float)lp;
}

// ---- reward and termination (rlmpc2.py:598-601, 701-735) ----
struct RewardArgs {
    int B;
    const double *state, *target, *control, *in_contact;   // [B,8] [B,8] [B,2] [B] (nullable: in contact)
    double* prev_cmd;                                        // [B,2] in/out
    const float* action;                                     // [B,34] raw action
    int32_t* episode_step;                                   // [B] in/out
    double* time_penalty;                                    // [B] in/out
    float *reward, *done;                                    // [B]
    dart_ppo_reward_cfg c;
};
__global__ void ppo_reward_kernel(const RewardArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const double* s = a.state + (size_t)b * 8;
    const double* tg = a.target + (size_t)b * 8;
    const double ex = fabs(tg[0] - s[0]), ey = fabs(tg[2] - s[2]);
    const double pos_err = sqrt(ex * ex + ey * ey), vel_err = sqrt(s[1] * s[1] + s[3] * s[3]);
    double n2 = 0.0;
    const double dz = a.c.max_delta * a.c.action_scale;
    for (int j = 0; j < ACT; ++j) { const double d = (double)(a.action[(size_t)b * ACT + j] * (float)dz); n2 += d * d; }
    double change = sqrt(n2);
    const double rms = change / sqrt((double)ACT);
    if (rms > a.c.max_per_dim_rms) change *= (double)(float)(a.c.max_per_dim_rms / (rms + 1e-12));
    const double c0 = a.control[2 * b], c1 = a.control[2 * b + 1];
    const double rate = fabs(c0 - a.prev_cmd[2 * b]) + fabs(c1 - a.prev_cmd[2 * b + 1]);
    a.prev_cmd[2 * b] = c0; a.prev_cmd[2 * b + 1] = c1;
    const double pt = exp(-(pos_err * pos_err) / (2.0 * a.c.sigma_pos * a.c.sigma_pos));
    const double vt = exp(-(vel_err * vel_err) / (2.0 * a.c.sigma_vel * a.c.sigma_vel));
    const double tp = a.time_penalty[b];
    double r = a.c.w_pos * pt + a.c.w_vel * pt * vt - a.c.w_change * change - a.c.w_d_ctrl * rate - tp;
    if (pos_err < a.c.success_tol && vel_err < a.c.success_tol) r += a.c.success_bonus;
    bool done = false;
    const int step = a.episode_step[b] + 1;
    if (fabs(s[0]) > a.c.tray_limit[0] || fabs(s[2]) > a.c.tray_limit[1]) { r -= a.c.oob_penalty; done = true; }
    if (a.in_contact && a.in_contact[b] == 0.0) r -= a.c.no_contact_penalty;
    if (step >= a.c.max_episode_steps) done = true;
    a.reward[b] = (float)r;
    a.done[b] = done ? 1.f : 0.f;
    a.episode_step[b] = done ? 0 : step;
    a.time_penalty[b] = done ? 0.0 : tp + a.c.time_penalty_inc;
}

// ---- GAE (rlmpc2.py:589-596), arrays [T,B], one thread per instance ----
__global__ void ppo_gae_kernel(int B, int T, const float* __restrict__ rew, const float* __restrict__ val,
                               const float* __restrict__ done, const float* __restrict__ last_value, double gamma,
                               double lam, float* __restrict__ adv, float* __restrict__ ret) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double gae = 0.0, vnext = (double)last_value[b];
    for (int t = T - 1; t >= 0; --t) {
        const size_t i = (size_t)t * B + b;
        const double nd = 1.0 - (double)done[i], v = (double)val[i];
        const double delta = (double)rew[i] + gamma * vnext * nd - v;
        gae = delta + gamma * lam * nd * gae;
        adv[i] = (float)gae;
        ret[i] = (float)(gae + v);
        vnext = v;
    }
}

__device__ double block_sum(double v, double* sh) {   // fixed-order tree: the same result on every launch
    const int t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (t < s) sh[t] += sh[t + s];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

// ---- in-place (x - mean) / (std + 1e-8), std with `ddof` delta degrees of freedom; single CTA ----
__global__ void __launch_bounds__(1024) ppo_normalize_kernel(long n, float* x, int ddof) {
    __shared__ double sh[1024];
    double s = 0.0;
    for (long i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i];
    const double mean = block_sum(s, sh) / (double)n;
    double q = 0.0;
    for (long i = threadIdx.x; i < n; i += blockDim.x) { const double d = (double)x[i] - mean; q += d * d; }
    const double denom = (double)(n - ddof);
    const double sd = sqrt(block_sum(q, sh) / (denom > 0.0 ? denom : 1.0));
    const double inv = 1.0 / (sd + 1e-8);
    for (long i = threadIdx.x; i < n; i += blockDim.x) x[i] = (float)(((double)x[i] - mean) * inv);
}

// ---- minibatch gather ----
struct GatherArgs {
    int M;
    const int64_t* idx;
    const float *obs, *act, *logp, *adv, *ret;
    float *o_obs, *o_act, *o_logp, *o_adv, *o_ret;
};
__global__ void ppo_gather_kernel(const GatherArgs a) {
    const int m = blockIdx.x;
    const int64_t src = a.idx[m];
    const float4* s4 = reinterpret_cast<const float4*>(a.obs + (size_t)src * OBS);   // 2080-byte rows: 16-byte aligned
    float4* d4 = reinterpret_cast<float4*>(a.o_obs + (size_t)m * OBS);
    for (int i = threadIdx.x; i < OBS / 4; i += blockDim.x) d4[i] = s4[i];
    for (int i = threadIdx.x; i < ACT; i += blockDim.x) a.o_act[(size_t)m * ACT + i] = a.act[(size_t)src * ACT + i];
    if (threadIdx.x == 0) { a.o_logp[m] = a.logp[src]; a.o_adv[m] = a.adv[src]; a.o_ret[m] = a.ret[src]; }
}

// ---- loss and its gradient w.r.t. the network outputs (rlmpc2.py:802-813) ----
struct LossArgs {
    int M;
    const float *mean, *value, *act, *old_logp, *adv, *ret, *log_std;
    float ls_min, ls_max, clip_eps, vf_coef;
    float *dmean, *dvalue;
    double* part;                      // [nblocks, LOSS_W]
};
__global__ void __launch_bounds__(LOSS_T) ppo_loss_kernel(const LossArgs a) {
    __shared__ float s_sd[ACT], s_in[ACT], s_ls[ACT];
    if (threadIdx.x < ACT) {
        const float raw = a.log_std[threadIdx.x];
        const float ls = fminf(fmaxf(raw, a.ls_min), a.ls_max);
        s_ls[threadIdx.x] = ls;
        s_sd[threadIdx.x] = fmaxf(expf(ls), 1e-6f);
        s_in[threadIdx.x] = (raw >= a.ls_min && raw <= a.ls_max) ? 1.f : 0.f;    // clamp passes the gradient inside its range
    }
    __syncthreads();
    const int m = blockIdx.x * LOSS_T + threadIdx.x;
    const bool on = m < a.M;
    double pl = 0.0, vl = 0.0, g = 0.0;
    float z[ACT];
    if (on) {
        double lp = 0.0;
#pragma unroll
        for (int j = 0; j < ACT; ++j) {
            const float d = a.act[(size_t)m * ACT + j] - a.mean[(size_t)m * ACT + j];
            z[j] = d / s_sd[j];
            lp += -0.5 * (double)z[j] * (double)z[j] - (double)s_ls[j] - HALF_LOG_2PI;
        }
        const double A = (double)a.adv[m];
        const double ratio = exp((double)(float)lp - (double)a.old_logp[m]);
        const double lo = 1.0 - (double)a.clip_eps, hi = 1.0 + (double)a.clip_eps;
        const double s1 = ratio * A, s2 = fmin(fmax(ratio, lo), hi) * A;
        pl = -fmin(s1, s2);
        g = (s1 <= s2) ? -A * ratio / (double)a.M : 0.0;        // dL/dlogp (torch.min ties split, both halves reach ratio)
        const double dv = (double)a.value[m] - (double)a.ret[m];
        vl = dv * dv;
        a.dvalue[m] = (float)((double)a.vf_coef * 2.0 * dv / (double)a.M);
#pragma unroll
        for (int j = 0; j < ACT; ++j) a.dmean[(size_t)m * ACT + j] = (float)(g * (double)z[j] / (double)s_sd[j]);
    }
    // block reduction of the 3 + 34 per-sample terms: shuffle tree inside each warp, then the 4 warp sums in a fixed order
    __shared__ double wsum[LOSS_T / 32][LOSS_W];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int q = 0; q < LOSS_W; ++q) {
        double v;
        if (q == 0) v = pl;
        else if (q == 1) v = vl;
        else if (q == 2) v = 0.0;
        else { const double zj = on ? (double)z[q - 3] : 0.0; v = on ? g * (zj * zj - 1.0) * (double)s_in[q - 3] : 0.0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) wsum[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < LOSS_W) {
        double r = 0.0;
#pragma unroll
        for (int w = 0; w < LOSS_T / 32; ++w) r += wsum[w][threadIdx.x];
        a.part[(size_t)blockIdx.x * LOSS_W + threadIdx.x] = r;
    }
}

// ---- layers 2-3, loss and their gradients for a tile of 64 samples in ONE kernel ----
// Everything between the two layer-1 GEMMs: h2 = tanh(h1 W2' + b2), mean / value, the per-sample loss terms, dmean / dvalue,
// dz2 = (dOut W3)(1 - h2^2), dz1 = (dz2 W2)(1 - h1^2) (written for the layer-1 weight-gradient GEMM) and this tile's
// partial weight/bias gradients of layers 2-3 (one row of part_small per CTA, summed in CTA order by the reduce kernel).
// The h1 tile, both weight orientations and all intermediates stay in shared memory (189 kB, one CTA per SM); replaces six
// grouped-GEMM launches + the loss kernel and the [M,128] h2 / dz2 round trips through HBM.  backward = 0 is the
// rollout-time forward (mean and value only).
constexpr int MS = 64, LDH = 132, LDW = 68, LDO = 36;
constexpr int SM_H1 = 0, SM_H2 = SM_H1 + MS * LDH, SM_D2 = SM_H2 + MS * LDH, SM_W2 = SM_D2 + MS * LDH,
              SM_W2T = SM_W2 + H2W * LDW, SM_W3 = SM_W2T + HID * LDH, SM_OUT = SM_W3 + 36 * LDW, SM_B = SM_OUT + MS * LDO,
              SM_FLOATS = SM_B + H2W + 36 + 3 * 36;
constexpr int MID_SMEM = SM_FLOATS * 4 + (GT / 32) * LOSS_W * 8;
// offsets inside a part_small row (flat layout relative to NPB)
constexpr int RS_W2A = OFF_W2A - NPB, RS_B2A = OFF_B2A - NPB, RS_W2C = OFF_W2C - NPB, RS_B2C = OFF_B2C - NPB,
              RS_W3A = OFF_W3A - NPB, RS_B3A = OFF_B3A - NPB, RS_W3C = OFF_W3C - NPB, RS_B3C = OFF_B3C - NPB;
struct MidArgs {
    int M, backward;
    int h1_splits;           // > 0: h1 points to `h1_splits` K-split partial pre-activations [split][M][128]; bias + tanh applied here
    const float *h1, *P, *act, *old_logp, *adv, *ret;
    const int64_t* idx;      // minibatch rows of act / old_logp / adv / ret (nullptr = rows 0..M-1)
    float ls_min, ls_max, clip_eps, vf_coef;
    float *mean, *value, *dz1, *part_small;
    double* loss_part;
};

__global__ void __launch_bounds__(GT, 1) ppo_mid_kernel(const MidArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* sH1 = sm + SM_H1; float* sH2 = sm + SM_H2; float* sD2 = sm + SM_D2; float* sW2 = sm + SM_W2;
    float* sW2T = sm + SM_W2T; float* sW3 = sm + SM_W3; float* sOut = sm + SM_OUT; float* sB2 = sm + SM_B;
    float* sB3 = sB2 + H2W; float* s_sd = sB3 + 36; float* s_ls = s_sd + 36; float* s_in = s_ls + 36;
    double* sLoss = reinterpret_cast<double*>(sm + SM_FLOATS);            // [8 warps][LOSS_W]
    const int t = threadIdx.x, ty = t / 16, tx = t % 16;
    const int s0 = blockIdx.x * MS;
    const float* P = a.P;
    // ---- P0: stage the h1 tile (rows past M zero-filled), W2 in both orientations, W3, biases, std ----
    for (int e = t; e < MS * (H2W / 4); e += GT) {
        const int s = e / (H2W / 4), c4 = (e % (H2W / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s0 + s < a.M) {
            v = *reinterpret_cast<const float4*>(a.h1 + (size_t)(s0 + s) * H2W + c4);
            if (a.h1_splits > 0) {            // small minibatches: the layer-1 forward was split over K (fixed summation order)
                for (int sp = 1; sp < a.h1_splits; ++sp) {
                    const float4 q = *reinterpret_cast<const float4*>(a.h1 + ((size_t)sp * a.M + s0 + s) * H2W + c4);
                    v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
                }
                const float4 b = *reinterpret_cast<const float4*>(P + OFF_B1 + c4);
                v.x = tanhf(v.x + b.x); v.y = tanhf(v.y + b.y); v.z = tanhf(v.z + b.z); v.w = tanhf(v.w + b.w);
            }
        }
        *reinterpret_cast<float4*>(&sH1[s * LDH + c4]) = v;
    }
    for (int e = t; e < H2W * HID; e += GT) {                              // e = c*64 + k over [W2A; W2C]
        const int c = e / HID, k = e % HID;
        const float w = (c < HID) ? P[OFF_W2A + c * HID + k] : P[OFF_W2C + (c - HID) * HID + k];
        sW2[c * LDW + k] = w;
        sW2T[k * LDH + c] = w;
    }
    for (int e = t; e < 36 * HID; e += GT) {
        const int j = e / HID, k = e % HID;
        sW3[j * LDW + k] = (j < ACT) ? P[OFF_W3A + j * HID + k] : (j == ACT ? P[OFF_W3C + k] : 0.f);
    }
    if (t < H2W) sB2[t] = (t < HID) ? P[OFF_B2A + t] : P[OFF_B2C + t - HID];
    if (t < 36) {
        sB3[t] = (t < ACT) ? P[OFF_B3A + t] : (t == ACT ? P[OFF_B3C] : 0.f);
        const float raw = (t < ACT) ? P[OFF_LS + t] : 0.f;
        const float ls = fminf(fmaxf(raw, a.ls_min), a.ls_max);
        s_ls[t] = ls; s_sd[t] = fmaxf(expf(ls), 1e-6f);
        s_in[t] = (raw >= a.ls_min && raw <= a.ls_max) ? 1.f : 0.f;
    }
    __syncthreads();
    // ---- P1: h2[s][c] = tanh(b2[c] + sum_k h1[s][k] W2[c][k]); thread: samples ty*4+i, columns tx*4+j of both networks ----
    {
        float acA[4][4], acC[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acA[i][j] = 0.f; acC[i][j] = 0.f; }
#pragma unroll 4
        for (int k = 0; k < HID; ++k) {
            const float4 wa = *reinterpret_cast<const float4*>(&sW2T[k * LDH + tx * 4]);
            const float4 wc = *reinterpret_cast<const float4*>(&sW2T[k * LDH + HID + tx * 4]);
            const float wav[4] = {wa.x, wa.y, wa.z, wa.w}, wcv[4] = {wc.x, wc.y, wc.z, wc.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float ha = sH1[(ty * 4 + i) * LDH + k], hc = sH1[(ty * 4 + i) * LDH + HID + k];
#pragma unroll
                for (int j = 0; j < 4; ++j) { acA[i][j] = fmaf(ha, wav[j], acA[i][j]); acC[i][j] = fmaf(hc, wcv[j], acC[i][j]); }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sH2[(ty * 4 + i) * LDH + tx * 4 + j] = tanhf(acA[i][j] + sB2[tx * 4 + j]);
                sH2[(ty * 4 + i) * LDH + HID + tx * 4 + j] = tanhf(acC[i][j] + sB2[HID + tx * 4 + j]);
            }
    }
    __syncthreads();
    // ---- P2: mean[s][j] (j < 34) and value[s] (j = 34); thread: sample t/4, columns (t%4) + 4i ----
    {
        const int s = t / 4, q = t % 4;
        float o[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = 0.f;
        for (int k = 0; k < HID; ++k) {
            const float ha = sH2[s * LDH + k], hc = sH2[s * LDH + HID + k];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const int j = q + 4 * i;
                o[i] = fmaf(j == ACT ? hc : ha, sW3[j * LDW + k], o[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int j = q + 4 * i;
            const float v = o[i] + sB3[j];
            sOut[s * LDO + j] = v;
            if (s0 + s < a.M) {
                if (j < ACT) a.mean[(size_t)(s0 + s) * ACT + j] = v;
                else if (j == ACT) a.value[s0 + s] = v;
            }
        }
    }
    if (!a.backward) return;
    __syncthreads();
    // ---- P3: per-sample loss terms and d(loss)/d(mean, value); four lanes per sample, lane q owns the action dimensions
    //      q, q+4, ...; sOut becomes dOut.  (One thread per sample left six of the eight warps waiting at the barrier below for
    //      29 % of the kernel's stall samples.) ----
    {
        const int s = t / 4, q = t % 4, m = s0 + s;
        const bool on = m < a.M;
        const size_t mr = on ? (a.idx ? (size_t)a.idx[m] : (size_t)m) : 0;      // row of the pooled rollout
        float zq[9];
        double lp = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int j = q + 4 * i;
            float zz = 0.f;
            if (on && j < ACT) {
                const float d = a.act[mr * ACT + j] - sOut[s * LDO + j];
                zz = d / s_sd[j];
                lp += -0.5 * (double)zz * (double)zz - (double)s_ls[j] - HALF_LOG_2PI;
            }
            zq[i] = zz;
        }
        lp += __shfl_xor_sync(0xffffffffu, lp, 1);
        lp += __shfl_xor_sync(0xffffffffu, lp, 2);
        double pl = 0.0, vl = 0.0, g = 0.0, dvs = 0.0;
        if (on) {
            const double A = (double)a.adv[mr];
            const double ratio = exp((double)(float)lp - (double)a.old_logp[mr]);
            const double lo = 1.0 - (double)a.clip_eps, hi = 1.0 + (double)a.clip_eps;
            const double s1 = ratio * A, s2 = fmin(fmax(ratio, lo), hi) * A;
            pl = -fmin(s1, s2);
            g = (s1 <= s2) ? -A * ratio / (double)a.M : 0.0;
            const double dv = (double)sOut[s * LDO + ACT] - (double)a.ret[mr];
            vl = dv * dv;
            dvs = (double)a.vf_coef * 2.0 * dv / (double)a.M;
        }
        __syncwarp();                                   // every lane of the sample has read mean / value before they are overwritten
        double part[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int j = q + 4 * i;
            part[i] = 0.0;
            if (j < ACT) {
                sOut[s * LDO + j] = (float)(g * (double)zq[i] / (double)s_sd[j]);
                part[i] = g * ((double)zq[i] * (double)zq[i] - 1.0) * (double)s_in[j];
            } else if (j == ACT) sOut[s * LDO + j] = (float)dvs;
        }
        if (q != 0) { pl = 0.0; vl = 0.0; }
        // sums over the 8 samples of the warp (lanes with equal q), then the 8 warp sums in a fixed order
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            pl += __shfl_xor_sync(0xffffffffu, pl, o);
            vl += __shfl_xor_sync(0xffffffffu, vl, o);
#pragma unroll
            for (int i = 0; i < 9; ++i) part[i] += __shfl_xor_sync(0xffffffffu, part[i], o);
        }
        const int lane = t & 31, warp = t >> 5;
        if (lane < 4) {
            if (lane == 0) { sLoss[warp * LOSS_W + 0] = pl; sLoss[warp * LOSS_W + 1] = vl; sLoss[warp * LOSS_W + 2] = 0.0; }
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const int j = q + 4 * i;
                if (j < ACT) sLoss[warp * LOSS_W + 3 + j] = part[i];
            }
        }
    }
    __syncthreads();
    if (t < LOSS_W) {
        double r = 0.0;
#pragma unroll
        for (int w = 0; w < GT / 32; ++w) r += sLoss[w * LOSS_W + t];
        a.loss_part[(size_t)blockIdx.x * LOSS_W + t] = r;
    }
    // ---- P4: dz2[s][c] = (sum_j dOut[s][j] W3[j][c]) (1 - h2^2) ----
    {
        float acA[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acA[i][j] = 0.f;
        for (int j = 0; j < ACT; ++j) {
            const float4 w = *reinterpret_cast<const float4*>(&sW3[j * LDW + tx * 4]);
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float d = sOut[(ty * 4 + i) * LDO + j];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acA[i][jj] = fmaf(d, wv[jj], acA[i][jj]);
            }
        }
        const float4 wc4 = *reinterpret_cast<const float4*>(&sW3[ACT * LDW + tx * 4]);
        const float wcv[4] = {wc4.x, wc4.y, wc4.z, wc4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float dvv = sOut[(ty * 4 + i) * LDO + ACT];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float ha = sH2[(ty * 4 + i) * LDH + tx * 4 + j], hc = sH2[(ty * 4 + i) * LDH + HID + tx * 4 + j];
                sD2[(ty * 4 + i) * LDH + tx * 4 + j] = acA[i][j] * (1.f - ha * ha);
                sD2[(ty * 4 + i) * LDH + HID + tx * 4 + j] = dvv * wcv[j] * (1.f - hc * hc);
            }
        }
    }
    __syncthreads();
    // ---- P5: dz1[s][k] = (sum_c dz2[s][c] W2[c][k]) (1 - h1^2)  -> HBM ----
    {
        float acA[4][4], acC[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acA[i][j] = 0.f; acC[i][j] = 0.f; }
#pragma unroll 4
        for (int c = 0; c < HID; ++c) {
            const float4 wa = *reinterpret_cast<const float4*>(&sW2[c * LDW + tx * 4]);
            const float4 wc = *reinterpret_cast<const float4*>(&sW2[(HID + c) * LDW + tx * 4]);
            const float wav[4] = {wa.x, wa.y, wa.z, wa.w}, wcv[4] = {wc.x, wc.y, wc.z, wc.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float da = sD2[(ty * 4 + i) * LDH + c], dc = sD2[(ty * 4 + i) * LDH + HID + c];
#pragma unroll
                for (int j = 0; j < 4; ++j) { acA[i][j] = fmaf(da, wav[j], acA[i][j]); acC[i][j] = fmaf(dc, wcv[j], acC[i][j]); }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int s = ty * 4 + i;
            if (s0 + s >= a.M) continue;
            float oa[4], oc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float ha = sH1[s * LDH + tx * 4 + j], hc = sH1[s * LDH + HID + tx * 4 + j];
                oa[j] = acA[i][j] * (1.f - ha * ha);
                oc[j] = acC[i][j] * (1.f - hc * hc);
            }
            *reinterpret_cast<float4*>(a.dz1 + (size_t)(s0 + s) * H2W + tx * 4) = make_float4(oa[0], oa[1], oa[2], oa[3]);
            *reinterpret_cast<float4*>(a.dz1 + (size_t)(s0 + s) * H2W + HID + tx * 4) = make_float4(oc[0], oc[1], oc[2], oc[3]);
        }
    }
    // ---- P6: this tile's weight / bias gradient partials of layers 2-3 ----
    float* row = a.part_small + (size_t)blockIdx.x * NPS_LD;
    {   // dW2[c][k] = sum_s dz2[s][c] h1[s][k]; thread: c = ty*4+i, k = tx*4+j, both networks
        float acA[4][4], acC[4][4], ba[4] = {0.f, 0.f, 0.f, 0.f}, bc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acA[i][j] = 0.f; acC[i][j] = 0.f; }
#pragma unroll 2
        for (int s = 0; s < MS; ++s) {
            const float4 da = *reinterpret_cast<const float4*>(&sD2[s * LDH + ty * 4]);
            const float4 dc = *reinterpret_cast<const float4*>(&sD2[s * LDH + HID + ty * 4]);
            const float4 ha = *reinterpret_cast<const float4*>(&sH1[s * LDH + tx * 4]);
            const float4 hc = *reinterpret_cast<const float4*>(&sH1[s * LDH + HID + tx * 4]);
            const float dav[4] = {da.x, da.y, da.z, da.w}, dcv[4] = {dc.x, dc.y, dc.z, dc.w};
            const float hav[4] = {ha.x, ha.y, ha.z, ha.w}, hcv[4] = {hc.x, hc.y, hc.z, hc.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ba[i] += dav[i]; bc[i] += dcv[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) { acA[i][j] = fmaf(dav[i], hav[j], acA[i][j]); acC[i][j] = fmaf(dcv[i], hcv[j], acC[i][j]); }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = ty * 4 + i;
            *reinterpret_cast<float4*>(row + RS_W2A + c * HID + tx * 4) = make_float4(acA[i][0], acA[i][1], acA[i][2], acA[i][3]);
            *reinterpret_cast<float4*>(row + RS_W2C + c * HID + tx * 4) = make_float4(acC[i][0], acC[i][1], acC[i][2], acC[i][3]);
            if (tx == 0) { row[RS_B2A + c] = ba[i]; row[RS_B2C + c] = bc[i]; }
        }
    }
    {   // dW3[j][k] = sum_s dOut[s][j] h2[s][k]; thread: k = t%64, rows j = t/64 + 4i (j = 34: the critic's row on its own h2 half)
        const int k = t % HID, q = t / HID;
        float o[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = 0.f;
        for (int s = 0; s < MS; ++s) {
            const float ha = sH2[s * LDH + k], hc = sH2[s * LDH + HID + k];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const int j = q + 4 * i;
                o[i] = fmaf(sOut[s * LDO + j], j == ACT ? hc : ha, o[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int j = q + 4 * i;
            if (j < ACT) row[RS_W3A + j * HID + k] = o[i];
            else if (j == ACT) row[RS_W3C + k] = o[i];
        }
        if (t <= ACT) {
            float b = 0.f;
            for (int s = 0; s < MS; ++s) b += sOut[s * LDO + t];
            if (t < ACT) row[RS_B3A + t] = b; else row[RS_B3C] = b;
        }
    }
}

// ---- sum the split partials into the flat gradient, per-block sum of squares for the global norm ----
struct ReduceArgs {
    int splits, splits_small, loss_blocks, M;
    const float* part;       // [splits, NPB]
    const float* part_small; // [splits_small, NPS_LD]
    const double* loss_part; // [loss_blocks, LOSS_W]
    const float* log_std;
    float ls_min, ls_max, ent_coef, vf_coef;
    float* grad;             // [NP]
    double* normpart;        // [gridDim.x]
    float* stats;            // [4] policy loss, value loss, entropy, (grad norm: written by the Adam kernel)
    long long* step_inc;     // optimiser step counter on the device, advanced here when an Adam launch follows (else nullptr)
};
// 64 parameters per block, four row groups per parameter (thread = (parameter, row group)): each thread sums every fourth partial
// row, the four group sums are combined in a fixed order.  (One thread per parameter walked up to 256 rows in one dependent chain.)
constexpr int RED_P = 64, RED_G = 4;
__global__ void __launch_bounds__(256) ppo_grad_reduce_kernel(const ReduceArgs a) {
    __shared__ double sh[256];
    __shared__ double sg[RED_G][RED_P];
    const int tx = threadIdx.x % RED_P, rg = threadIdx.x / RED_P;
    const int j = blockIdx.x * RED_P + tx;
    double acc = 0.0;
    if (j < NPB) {
        float f = 0.f;
        for (int s = rg; s < a.splits; s += RED_G) f += a.part[(size_t)s * NPB + j];
        acc = (double)f;
    } else if (j < NPW) {
        float f = 0.f;
        for (int s = rg; s < a.splits_small; s += RED_G) f += a.part_small[(size_t)s * NPS_LD + (j - NPB)];
        acc = (double)f;
    } else if (j < NP) {
        for (int b = rg; b < a.loss_blocks; b += RED_G) acc += a.loss_part[(size_t)b * LOSS_W + 3 + (j - NPW)];
    }
    sg[rg][tx] = acc;
    __syncthreads();
    float g = 0.f;
    if (rg == 0 && j < NP) {
        if (j < NPW) g = ((float)sg[0][tx] + (float)sg[1][tx]) + ((float)sg[2][tx] + (float)sg[3][tx]);
        else {
            double d = (sg[0][tx] + sg[1][tx]) + (sg[2][tx] + sg[3][tx]);
            const float raw = a.log_std[j - NPW];
            if (raw >= a.ls_min && raw <= a.ls_max) d -= (double)a.ent_coef;     // d(-ent_coef * entropy)/dlog_std
            g = (float)d;
        }
        a.grad[j] = g;
    }
    const double ss = block_sum((double)g * (double)g, sh);
    if (threadIdx.x == 0) a.normpart[blockIdx.x] = ss;
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.step_inc) *a.step_inc += 1;
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.stats) {
        double pl = 0.0, vl = 0.0, ent = 0.0;
        for (int b = 0; b < a.loss_blocks; ++b) { pl += a.loss_part[(size_t)b * LOSS_W]; vl += a.loss_part[(size_t)b * LOSS_W + 1]; }
        for (int q = 0; q < ACT; ++q) ent += 0.5 + HALF_LOG_2PI + (double)fminf(fmaxf(a.log_std[q], a.ls_min), a.ls_max);
        a.stats[0] = (float)(pl / a.M);
        a.stats[1] = (float)(vl / a.M);
        a.stats[2] = (float)ent;
    }
}

// ---- clip_grad_norm_ + Adam with L2 weight decay (torch.optim.Adam, rlmpc2.py:561, 816-817) ----
struct AdamArgs {
    int nparts;
    const double* normpart;
    const float* grad;
    float *param, *m, *v;
    float max_norm, wd, beta1, beta2, eps;
    double lr, beta1d, beta2d;
    const long long* step;   // device-resident step count (already advanced for this step): keeps the launch replayable from a CUDA graph
    float* stats;
};
__global__ void __launch_bounds__(256) ppo_adam_kernel(const AdamArgs a) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < a.nparts; i += 256) s += a.normpart[i];
    const float gnorm = (float)sqrt(block_sum(s, sh));
    const float coef = fminf(a.max_norm / (gnorm + 1e-6f), 1.f);
    __shared__ float s_step_size, s_bc2_sqrt;
    if (threadIdx.x == 0) {               // torch computes these in Python floats (double) and applies them in FP32
        const double st = (double)*a.step;
        s_step_size = (float)(a.lr / (1.0 - pow(a.beta1d, st)));
        s_bc2_sqrt = (float)sqrt(1.0 - pow(a.beta2d, st));
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.stats) a.stats[3] = gnorm;
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= NP) return;
    const float p = a.param[j];
    const float g = fmaf(a.wd, p, a.grad[j] * coef);
    const float m = a.m[j] + (g - a.m[j]) * (1.f - a.beta1);
    const float v = fmaf(1.f - a.beta2, g * g, a.v[j] * a.beta2);
    a.m[j] = m;
    a.v[j] = v;
    const float denom = sqrtf(v) / s_bc2_sqrt + a.eps;
    a.param[j] = p - s_step_size * (m / denom);
}


// ---- data-parallel training: the all-reduced gradient comes back from the host side's collective ----
__global__ void __launch_bounds__(256) ppo_grad_import_kernel(const float* __restrict__ src, float scale, float* __restrict__ grad,
                                                              double* __restrict__ normpart, long long* step_inc) {
    __shared__ double sh[256];
    const int j = blockIdx.x * 256 + threadIdx.x;
    float g = 0.f;
    if (j < NP) { g = src[j] * scale; grad[j] = g; }
    const double ss = block_sum((double)g * (double)g, sh);
    if (threadIdx.x == 0) normpart[blockIdx.x] = ss;
    if (blockIdx.x == 0 && threadIdx.x == 0 && step_inc) *step_inc += 1;
}

}  // namespace

struct dart_ppo {
    int device, capacity, splits_cap, loss_blocks_cap, fused, sms;
    int tc;                              // layer-1 GEMMs of large minibatches on tcgen05 (ppo_tc.cu); DART_PPO_SIMT=1 keeps the FP32 SIMT kernels
    dart_ppo_cfg cfg;
    float *param, *grad, *m, *v;         // [NP]
    float *h1, *h2, *dz1, *dz2;          // [capacity, 128]
    float* h1part;                       // K-split layer-1 pre-activations: [FWD_SPLITS, rows < BIG_MIN_ROWS, 128] or [BIG_FWD_SPLITS, capacity, 128]
    float *mean, *value, *dmean, *dvalue;
    float *mb_obs, *mb_act, *mb_logp, *mb_adv, *mb_ret;
    float *part, *part_small;            // [MAX_SPLITS, NPB], [MAX_SPLITS_SMALL, NPS]
    double *loss_part, *normpart;
    long long* step_dev;                 // optimiser step count, on the device
    int64_t launches;
};

namespace {

int launch_group(dart_ppo* h, GemmGroup& g, int maxM, int maxN, cudaStream_t st) {
    dim3 grid((maxN + BN - 1) / BN, (maxM + BM - 1) / BM, g.count * g.splits);
    gemm_kernel<<<grid, GT, 0, st>>>(g);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

int launch_big(dart_ppo* h, const GemmProb& p, int splits, int kchunk, long split_stride, cudaStream_t st) {
    dim3 grid((p.N + TN - 1) / TN, (p.M + TM - 1) / TM, splits);
    gemm128_kernel<<<grid, GT, 0, st>>>(p, splits, kchunk, split_stride);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

GemmProb fwd_prob(const float* X, int ldx, const float* W, const float* b, float* Y, int ldy, int M, int N, int K, int mode) {
    GemmProb p;
    memset(&p, 0, sizeof(p));
    p.A = X; p.lda = ldx; p.ta = 0;
    p.B = W; p.ldb = K; p.tb = 1;          // B(k,n) = W[n*K + k]
    p.C = Y; p.ldc = ldy; p.M = M; p.N = N; p.K = K; p.mode = mode; p.bias = b;
    return p;
}

// Both networks' forward pass for M rows of `obs`: h1, h2 [M,128] (actor columns 0-63, critic 64-127), mean [M,34], value [M].
int forward(dart_ppo* h, int M, const float* obs, cudaStream_t st, bool with_tail = true, const int64_t* gidx = nullptr) {
    const float* P = h->param;
    GemmGroup g;
    memset(&g, 0, sizeof(g));
    g.count = 1; g.splits = 1; g.kchunk = OBS; g.split_stride = 0;
    g.p[0] = fwd_prob(obs, OBS, P + OFF_W1, P + OFF_B1, h->h1, H2W, M, H2W, OBS, 1);
    if (gidx) { g.p[0].gidx = gidx; g.p[0].gmode = 1; }      // rows of the pooled rollout, gathered by the loaders
    int rc;
    if (M >= BIG_MIN_ROWS && h->fused && h->tc) {      // tensor cores: 3xTF32, one CTA per 128 rows, pre-activation to h1part
        rc = dart_ppo_tc::l1_forward(M, obs, gidx, P + OFF_W1, h->h1part, st);
        h->launches += 1;
    } else if (M >= BIG_MIN_ROWS && h->fused) {  // two K halves: 2 x (M/128) CTAs fill two slots per SM (one half alone leaves 128 CTAs on 148 SMs)
        g.p[0].mode = 0; g.p[0].C = h->h1part;
        rc = launch_big(h, g.p[0], BIG_FWD_SPLITS, BIG_FWD_KCHUNK, (long)M * H2W, st);
    } else if (M >= BIG_MIN_ROWS) rc = launch_big(h, g.p[0], 1, OBS, 0, st);
    else if (h->fused) {                  // few rows: split K over FWD_SPLITS CTAs per tile; bias + tanh move into the tile kernel
        g.splits = FWD_SPLITS; g.kchunk = FWD_KCHUNK; g.split_stride = (long)M * H2W;
        g.p[0].mode = 0; g.p[0].C = h->h1part;
        rc = launch_group(h, g, M, H2W, st);
        g.splits = 1; g.kchunk = OBS; g.split_stride = 0;
    } else rc = launch_group(h, g, M, H2W, st);
    if (rc != DART_OK) return rc;
    if (h->fused) {
        if (!with_tail) return DART_OK;
        MidArgs ma;
        memset(&ma, 0, sizeof(ma));
        ma.M = M; ma.backward = 0; ma.h1 = h->h1; ma.P = P; ma.mean = h->mean; ma.value = h->value;
        ma.h1 = h->h1part; ma.h1_splits = M < BIG_MIN_ROWS ? FWD_SPLITS : (h->tc ? 1 : BIG_FWD_SPLITS);
        ma.ls_min = (float)h->cfg.log_std_min; ma.ls_max = (float)h->cfg.log_std_max;
        ppo_mid_kernel<<<(M + MS - 1) / MS, GT, MID_SMEM, st>>>(ma);
        h->launches += 1;
        return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
    }
    g.count = 2; g.kchunk = HID;
    g.p[0] = fwd_prob(h->h1, H2W, P + OFF_W2A, P + OFF_B2A, h->h2, H2W, M, HID, HID, 1);
    g.p[1] = fwd_prob(h->h1 + HID, H2W, P + OFF_W2C, P + OFF_B2C, h->h2 + HID, H2W, M, HID, HID, 1);
    rc = launch_group(h, g, M, HID, st);
    if (rc != DART_OK) return rc;
    g.p[0] = fwd_prob(h->h2, H2W, P + OFF_W3A, P + OFF_B3A, h->mean, ACT, M, ACT, HID, 2);
    g.p[1] = fwd_prob(h->h2 + HID, H2W, P + OFF_W3C, P + OFF_B3C, h->value, 1, M, 1, HID, 2);
    return launch_group(h, g, M, ACT, st);
}

// dW[N,K] (+ db[N]) = dY^T X over the minibatch, split into g.splits chunks of samples
GemmProb wgrad_prob(const float* dY, int lddy, const float* X, int ldx, float* part, int off, int M, int N, int K) {
    GemmProb p;
    memset(&p, 0, sizeof(p));
    p.A = dY; p.lda = lddy; p.ta = 1;      // A(n, sample) = dY[sample*lddy + n]
    p.B = X; p.ldb = ldx; p.tb = 0;        // B(sample, k) = X[sample*ldx + k]
    p.C = part + off; p.ldc = K; p.M = N; p.N = K; p.K = M; p.mode = 0;
    p.rowsum = part + off + N * K;          // the bias follows its weight matrix in the flat layout
    return p;
}

// dX[M,K] = (dY[M,N] W[N,K]) * (1 - Hprev^2)
GemmProb dgrad_prob(const float* dY, int lddy, const float* W, float* dX, int lddx, const float* Hprev, int ldh, int M, int N, int K) {
    GemmProb p;
    memset(&p, 0, sizeof(p));
    p.A = dY; p.lda = lddy; p.ta = 0;
    p.B = W; p.ldb = K; p.tb = 0;          // B(n,k) = W[n*K + k]
    p.C = dX; p.ldc = lddx; p.M = M; p.N = K; p.K = N; p.mode = 3; p.H = Hprev; p.ldh = ldh;
    return p;
}

// The first (unfused) kernel chain for layers 2-3 and the loss: kept as the A/B reference of ppo_mid_kernel (DART_PPO_UNFUSED=1).
int backward_unfused(dart_ppo* h, int M, const float* act, const float* old_logp, const float* adv, const float* ret,
                     cudaStream_t st, int* loss_blocks_out, int* splits_small_out) {
    int rc;
    const dart_ppo_cfg& c = h->cfg;
    const float* P = h->param;
    const int loss_blocks = (M + LOSS_T - 1) / LOSS_T;
    LossArgs la;
    la.M = M; la.mean = h->mean; la.value = h->value; la.act = act; la.old_logp = old_logp; la.adv = adv; la.ret = ret;
    la.log_std = P + OFF_LS; la.ls_min = (float)c.log_std_min; la.ls_max = (float)c.log_std_max;
    la.clip_eps = (float)c.clip_eps; la.vf_coef = (float)c.vf_coef; la.dmean = h->dmean; la.dvalue = h->dvalue;
    la.part = h->loss_part;
    ppo_loss_kernel<<<loss_blocks, LOSS_T, 0, st>>>(la);
    h->launches += 1;
    if (cudaGetLastError() != cudaSuccess) return DART_ERR_CUDA;

    // backward: weight gradients split over the minibatch, data gradients with the fused tanh'
    int splits_small = (M + 63) / 64;
    if (splits_small > MAX_SPLITS_SMALL) splits_small = MAX_SPLITS_SMALL;
    int kchunk_small = (M + splits_small - 1) / splits_small;
    kchunk_small = (kchunk_small + BK - 1) / BK * BK;
    float* ps = h->part_small;                        // layers 2-3 block: flat offsets relative to NPB
    GemmGroup gw;
    memset(&gw, 0, sizeof(gw));
    gw.count = 2; gw.splits = splits_small; gw.kchunk = kchunk_small; gw.split_stride = NPS_LD;
    GemmGroup gd;
    memset(&gd, 0, sizeof(gd));
    gd.count = 2; gd.splits = 1; gd.split_stride = 0;
    // layer 3
    gw.p[0] = wgrad_prob(h->dmean, ACT, h->h2, H2W, ps, OFF_W3A - NPB, M, ACT, HID);
    gw.p[1] = wgrad_prob(h->dvalue, 1, h->h2 + HID, H2W, ps, OFF_W3C - NPB, M, 1, HID);
    if ((rc = launch_group(h, gw, ACT, HID, st)) != DART_OK) return rc;
    gd.kchunk = ACT;
    gd.p[0] = dgrad_prob(h->dmean, ACT, P + OFF_W3A, h->dz2, H2W, h->h2, H2W, M, ACT, HID);
    gd.p[1] = dgrad_prob(h->dvalue, 1, P + OFF_W3C, h->dz2 + HID, H2W, h->h2 + HID, H2W, M, 1, HID);
    if ((rc = launch_group(h, gd, M, HID, st)) != DART_OK) return rc;
    // layer 2
    gw.p[0] = wgrad_prob(h->dz2, H2W, h->h1, H2W, ps, OFF_W2A - NPB, M, HID, HID);
    gw.p[1] = wgrad_prob(h->dz2 + HID, H2W, h->h1 + HID, H2W, ps, OFF_W2C - NPB, M, HID, HID);
    if ((rc = launch_group(h, gw, HID, HID, st)) != DART_OK) return rc;
    gd.kchunk = HID;
    gd.p[0] = dgrad_prob(h->dz2, H2W, P + OFF_W2A, h->dz1, H2W, h->h1, H2W, M, HID, HID);
    gd.p[1] = dgrad_prob(h->dz2 + HID, H2W, P + OFF_W2C, h->dz1 + HID, H2W, h->h1 + HID, H2W, M, HID, HID);
    if ((rc = launch_group(h, gd, M, HID, st)) != DART_OK) return rc;
    *loss_blocks_out = loss_blocks; *splits_small_out = splits_small;
    return DART_OK;
}

void free_all(dart_ppo* h) {
    void* p[] = {h->param, h->grad, h->m, h->v, h->h1, h->h2, h->dz1, h->dz2, h->mean, h->value, h->dmean, h->dvalue,
                 h->mb_obs, h->mb_act, h->mb_logp, h->mb_adv, h->mb_ret, h->part, h->part_small, h->loss_part, h->normpart, h->step_dev, h->h1part};
    for (void* q : p) if (q) cudaFree(q);
}

}  // namespace

extern "C" int dart_ppo_default_cfg(dart_ppo_cfg* c) {
    if (!c) return DART_ERR_ARG;
    // rlmpc2.py:202-226 packet defaults, torch.optim.Adam defaults, Policy log_std range (:57-61)
    c->lr = 3e-4; c->weight_decay = 1e-5; c->beta1 = 0.9; c->beta2 = 0.999; c->adam_eps = 1e-8;
    c->clip_eps = 0.2; c->vf_coef = 0.25; c->ent_coef = 0.01; c->max_grad_norm = 0.5;
    c->log_std_min = log(1e-2); c->log_std_max = log(2.0);
    return DART_OK;
}

extern "C" int dart_ppo_default_reward_cfg(dart_ppo_reward_cfg* c) {
    if (!c) return DART_ERR_ARG;
    // rlmpc2.py:701-735 with the defaults of the packet.get(...) calls there
    c->max_delta = 0.1; c->action_scale = 1.0; c->max_per_dim_rms = 0.5;
    c->sigma_pos = 0.02; c->sigma_vel = 0.02; c->w_pos = 60.0; c->w_vel = 30.0; c->w_change = 1e-3; c->w_d_ctrl = 5.0;
    c->success_tol = 0.01; c->success_bonus = 20.0; c->oob_penalty = 20.0; c->no_contact_penalty = 10.0;
    c->tray_limit[0] = 0.2; c->tray_limit[1] = 0.15; c->max_episode_steps = 1000; c->time_penalty_inc = 1e-4;
    return DART_OK;
}

extern "C" int dart_ppo_nparams(void) { return NP; }

extern "C" int dart_ppo_destroy(dart_ppo_handle h) {
    if (!h) return DART_ERR_ARG;
    cudaSetDevice(h->device);
    free_all(h);
    delete h;
    return DART_OK;
}

extern "C" int dart_ppo_create(dart_ppo_handle* out, int device, int32_t obs_dim, int32_t hidden, int32_t act_dim,
                               int32_t capacity, const float* params_host, const dart_ppo_cfg* cfg) {
    if (!out || !params_host || !cfg || capacity < 1) return DART_ERR_ARG;
    if (obs_dim != OBS || hidden != HID || act_dim != ACT) return DART_ERR_UNSUPPORTED;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return DART_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) return DART_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return DART_ERR_CUDA;
    dart_ppo* h = new (std::nothrow) dart_ppo();
    if (!h) return DART_ERR_ALLOC;
    memset(h, 0, sizeof(*h));
    h->device = device; h->capacity = capacity; h->cfg = *cfg;
    h->loss_blocks_cap = (capacity + MS - 1) / MS;          // the fused path has one loss row per 64-sample tile
    cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device);
    h->fused = getenv("DART_PPO_UNFUSED") ? 0 : 1;
    h->tc = getenv("DART_PPO_SIMT") ? 0 : 1;         // A/B switch: the first, unfused kernel chain
    const size_t cap = (size_t)capacity;
    struct { void** p; size_t bytes; } al[] = {
        {(void**)&h->param, NP * sizeof(float)}, {(void**)&h->grad, NP * sizeof(float)},
        {(void**)&h->m, NP * sizeof(float)}, {(void**)&h->v, NP * sizeof(float)},
        {(void**)&h->h1, cap * H2W * sizeof(float)}, {(void**)&h->h2, cap * H2W * sizeof(float)},
        {(void**)&h->dz1, cap * H2W * sizeof(float)}, {(void**)&h->dz2, cap * H2W * sizeof(float)},
        {(void**)&h->h1part, (cap < BIG_MIN_ROWS ? (size_t)FWD_SPLITS * cap : (size_t)(BIG_FWD_SPLITS * cap > (size_t)FWD_SPLITS * BIG_MIN_ROWS ? BIG_FWD_SPLITS * cap : (size_t)FWD_SPLITS * BIG_MIN_ROWS)) * H2W * sizeof(float)},
        {(void**)&h->mean, cap * ACT * sizeof(float)}, {(void**)&h->value, cap * sizeof(float)},
        {(void**)&h->dmean, cap * ACT * sizeof(float)}, {(void**)&h->dvalue, cap * sizeof(float)},
        {(void**)&h->mb_obs, cap * OBS * sizeof(float)}, {(void**)&h->mb_act, cap * ACT * sizeof(float)},
        {(void**)&h->mb_logp, cap * sizeof(float)}, {(void**)&h->mb_adv, cap * sizeof(float)},
        {(void**)&h->mb_ret, cap * sizeof(float)}, {(void**)&h->part, (size_t)MAX_SPLITS * NPB * sizeof(float)},
        {(void**)&h->part_small, (size_t)(h->loss_blocks_cap > MAX_SPLITS_SMALL ? h->loss_blocks_cap : MAX_SPLITS_SMALL) * NPS_LD * sizeof(float)},
        {(void**)&h->loss_part, (size_t)h->loss_blocks_cap * LOSS_W * sizeof(double)},
        {(void**)&h->normpart, (size_t)((NP + RED_P - 1) / RED_P) * sizeof(double)}, {(void**)&h->step_dev, sizeof(long long)}};
    int rc = DART_OK;
    for (auto& a : al) {
        if (cudaMalloc(a.p, a.bytes) != cudaSuccess) { *a.p = nullptr; rc = DART_ERR_ALLOC; break; }
        if (cudaMemset(*a.p, 0, a.bytes) != cudaSuccess) { rc = DART_ERR_CUDA; break; }
    }
    if (rc == DART_OK && cudaMemcpy(h->param, params_host, NP * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
        rc = DART_ERR_CUDA;
    if (rc == DART_OK && cudaFuncSetAttribute(ppo_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MID_SMEM) != cudaSuccess)
        rc = DART_ERR_CUDA;
    if (rc != DART_OK) { cudaGetLastError(); free_all(h); delete h; return rc; }
    *out = h;
    return DART_OK;
}

extern "C" int dart_ppo_get_state(dart_ppo_handle h, float* params_host, float* m_host, float* v_host, int64_t* step) {
    if (!h) return DART_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return DART_ERR_CUDA;
    if (cudaDeviceSynchronize() != cudaSuccess) return DART_ERR_CUDA;
    const size_t n = NP * sizeof(float);
    if ((params_host && cudaMemcpy(params_host, h->param, n, cudaMemcpyDeviceToHost) != cudaSuccess) ||
        (m_host && cudaMemcpy(m_host, h->m, n, cudaMemcpyDeviceToHost) != cudaSuccess) ||
        (v_host && cudaMemcpy(v_host, h->v, n, cudaMemcpyDeviceToHost) != cudaSuccess))
        return DART_ERR_CUDA;
    long long st = 0;
    if (cudaMemcpy(&st, h->step_dev, sizeof(st), cudaMemcpyDeviceToHost) != cudaSuccess) return DART_ERR_CUDA;
    if (step) *step = (int64_t)st;
    return DART_OK;
}

extern "C" int dart_ppo_set_state(dart_ppo_handle h, const float* params_host, const float* m_host, const float* v_host,
                                  int64_t step) {
    if (!h || step < 0) return DART_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return DART_ERR_CUDA;
    if (cudaDeviceSynchronize() != cudaSuccess) return DART_ERR_CUDA;
    const size_t n = NP * sizeof(float);
    if ((params_host && cudaMemcpy(h->param, params_host, n, cudaMemcpyHostToDevice) != cudaSuccess) ||
        (m_host && cudaMemcpy(h->m, m_host, n, cudaMemcpyHostToDevice) != cudaSuccess) ||
        (v_host && cudaMemcpy(h->v, v_host, n, cudaMemcpyHostToDevice) != cudaSuccess))
        return DART_ERR_CUDA;
    const long long st = (long long)step;
    return cudaMemcpy(h->step_dev, &st, sizeof(st), cudaMemcpyHostToDevice) == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" int dart_ppo_get_grad(dart_ppo_handle h, float* grad_host) {
    if (!h || !grad_host) return DART_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return DART_ERR_CUDA;
    return cudaMemcpy(grad_host, h->grad, NP * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" const float* dart_ppo_params_dev(dart_ppo_handle h) { return h ? h->param : nullptr; }
extern "C" int64_t dart_ppo_launch_count(dart_ppo_handle h) { return h ? h->launches : -1; }

extern "C" int dart_ppo_act(dart_ppo_handle h, int32_t B, const float* obs, const float* eps, float* action, float* logp,
                            float* value, float* mean, void* stream) {
    if (!h || B < 0 || B > h->capacity || !obs || !action || !logp || !value) return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    { int cur = -1; if (cudaGetDevice(&cur) != cudaSuccess || cur != h->device) return DART_ERR_ARG; }   // launch from the handle's device
    cudaStream_t st = (cudaStream_t)stream;
    int rc = forward(h, B, obs, st);
    if (rc != DART_OK) return rc;
    ppo_sample_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, h->mean, h->param + OFF_LS, (float)h->cfg.log_std_min,
                                                      (float)h->cfg.log_std_max, eps, action, logp);
    h->launches += 1;
    if (cudaMemcpyAsync(value, h->value, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
        (mean && cudaMemcpyAsync(mean, h->mean, (size_t)B * ACT * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess))
        return DART_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" int dart_ppo_reward(int32_t B, const dart_ppo_reward_cfg* cfg, const double* state, const double* target,
                               const double* control, double* prev_cmd, const float* action, const double* in_contact,
                               int32_t* episode_step, double* time_penalty, float* reward, float* done, void* stream) {
    if (B < 0 || !cfg || !state || !target || !control || !prev_cmd || !action || !episode_step || !time_penalty || !reward || !done)
        return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    RewardArgs a;
    a.B = B; a.state = state; a.target = target; a.control = control; a.in_contact = in_contact; a.prev_cmd = prev_cmd;
    a.action = action; a.episode_step = episode_step; a.time_penalty = time_penalty; a.reward = reward; a.done = done;
    a.c = *cfg;
    ppo_reward_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" int dart_ppo_gae(int32_t B, int32_t T, const float* rewards, const float* values, const float* dones,
                            const float* last_value, double gamma, double lam, float* adv, float* ret, void* stream) {
    if (B < 0 || T < 0 || !rewards || !values || !dones || !last_value || !adv || !ret) return DART_ERR_ARG;
    if (B == 0 || T == 0) return DART_OK;
    ppo_gae_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, T, rewards, values, dones, last_value, gamma, lam, adv, ret);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" int dart_ppo_normalize(int64_t n, float* x, int32_t ddof, void* stream) {
    if (n < 0 || !x || ddof < 0 || ddof > 1) return DART_ERR_ARG;
    if (n == 0) return DART_OK;
    ppo_normalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((long)n, x, ddof);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" int dart_ppo_update(dart_ppo_handle h, int32_t M, const int64_t* idx, const float* obs, const float* act,
                               const float* old_logp, const float* adv, const float* ret, int32_t apply, float* stats,
                               void* stream) {
    if (!h || M < 1 || M > h->capacity || !obs || !act || !old_logp || !adv || !ret) return DART_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(obs) & 15) != 0) return DART_ERR_ARG;
    { int cur = -1; if (cudaGetDevice(&cur) != cudaSuccess || cur != h->device) return DART_ERR_ARG; }   // launch from the handle's device
    cudaStream_t st = (cudaStream_t)stream;
    const dart_ppo_cfg& c = h->cfg;
    const int64_t* gidx = nullptr;
    if (idx && h->fused) gidx = idx;                  // the fused path gathers inside its loaders: no minibatch copy
    else if (idx) {
        GatherArgs ga{M, idx, obs, act, old_logp, adv, ret, h->mb_obs, h->mb_act, h->mb_logp, h->mb_adv, h->mb_ret};
        ppo_gather_kernel<<<M, 128, 0, st>>>(ga);
        h->launches += 1;
        obs = h->mb_obs; act = h->mb_act; old_logp = h->mb_logp; adv = h->mb_adv; ret = h->mb_ret;
    }
    int rc = forward(h, M, obs, st, /*with_tail=*/!h->fused, gidx);
    if (rc != DART_OK) return rc;
    const float* P = h->param;
    int loss_blocks, splits, kchunk, splits_small;
    splits = (M + 255) / 256;
    int max_splits = MAX_SPLITS;
    if (M >= BIG_MIN_ROWS) {              // 5 column tiles x splits CTAs, two per SM: keep them to a single wave
        const int fit = (2 * h->sms) / ((OBS + TN - 1) / TN);
        if (fit >= 1 && fit < max_splits) max_splits = fit;
    }
    if (splits > max_splits) splits = max_splits;
    kchunk = (M + splits - 1) / splits;
    kchunk = (kchunk + BK - 1) / BK * BK;
    splits = (M + kchunk - 1) / kchunk;
    const float ls_min = (float)c.log_std_min, ls_max = (float)c.log_std_max;
    if (h->fused) {
        const int tiles = (M + MS - 1) / MS;
        MidArgs ma;
        memset(&ma, 0, sizeof(ma));
        ma.h1 = h->h1;
        ma.h1 = h->h1part; ma.h1_splits = M < BIG_MIN_ROWS ? FWD_SPLITS : (h->tc ? 1 : BIG_FWD_SPLITS);
        ma.M = M; ma.backward = 1; ma.P = P; ma.idx = gidx; ma.act = act; ma.old_logp = old_logp; ma.adv = adv; ma.ret = ret;
        ma.ls_min = ls_min; ma.ls_max = ls_max; ma.clip_eps = (float)c.clip_eps; ma.vf_coef = (float)c.vf_coef;
        ma.mean = h->mean; ma.value = h->value; ma.dz1 = h->dz1; ma.part_small = h->part_small; ma.loss_part = h->loss_part;
        ppo_mid_kernel<<<tiles, GT, MID_SMEM, st>>>(ma);
        h->launches += 1;
        if (cudaGetLastError() != cudaSuccess) return DART_ERR_CUDA;
        loss_blocks = tiles; splits_small = tiles;
    } else {
        rc = backward_unfused(h, M, act, old_logp, adv, ret, st, &loss_blocks, &splits_small);
        if (rc != DART_OK) return rc;
    }
    {   // layer 1 (both networks at once: dz1 is [M,128])
        GemmGroup gw;
        memset(&gw, 0, sizeof(gw));
        gw.count = 1; gw.splits = splits; gw.kchunk = kchunk; gw.split_stride = NPB;
        gw.p[0] = wgrad_prob(h->dz1, H2W, obs, OBS, h->part, OFF_W1, M, H2W, OBS);
        if (gidx) { gw.p[0].gidx = gidx; gw.p[0].gmode = 2; }
        if (M >= BIG_MIN_ROWS && h->fused && h->tc) {
            // tensor cores: four feature slabs x `splits` sample ranges, one CTA per SM in a single wave
            int per = 0;
            int want = h->sms / 4;                 // four feature slabs per split
            if (want > MAX_SPLITS) want = MAX_SPLITS;
            splits = dart_ppo_tc::l1_wgrad_splits(M, want, &per);
            rc = dart_ppo_tc::l1_wgrad(M, splits, per, h->dz1, obs, gidx, h->part + OFF_W1, NPB, st);
            h->launches += 1;
        } else
        rc = M >= BIG_MIN_ROWS ? launch_big(h, gw.p[0], splits, kchunk, NPB, st) : launch_group(h, gw, H2W, OBS, st);
        if (rc != DART_OK) return rc;
    }
    LossArgs la;            // (only the clamp bounds are read below)
    la.ls_min = ls_min; la.ls_max = ls_max; la.vf_coef = (float)c.vf_coef;

    const int nred = (NP + 255) / 256, nred_r = (NP + RED_P - 1) / RED_P;
    ReduceArgs ra;
    ra.splits = splits; ra.splits_small = splits_small; ra.loss_blocks = loss_blocks; ra.M = M; ra.part = h->part;
    ra.part_small = h->part_small; ra.loss_part = h->loss_part;
    ra.log_std = P + OFF_LS; ra.ls_min = la.ls_min; ra.ls_max = la.ls_max; ra.ent_coef = (float)c.ent_coef;
    ra.vf_coef = la.vf_coef; ra.grad = h->grad; ra.normpart = h->normpart; ra.stats = stats;
    ra.step_inc = apply ? h->step_dev : nullptr;
    ppo_grad_reduce_kernel<<<nred_r, 256, 0, st>>>(ra);
    h->launches += 1;
    if (cudaGetLastError() != cudaSuccess) return DART_ERR_CUDA;
    if (!apply) return DART_OK;

    AdamArgs aa;
    aa.nparts = nred_r; aa.normpart = h->normpart; aa.grad = h->grad; aa.param = h->param; aa.m = h->m; aa.v = h->v;
    aa.max_norm = (float)c.max_grad_norm; aa.wd = (float)c.weight_decay; aa.beta1 = (float)c.beta1; aa.beta2 = (float)c.beta2;
    aa.eps = (float)c.adam_eps; aa.lr = c.lr; aa.beta1d = c.beta1; aa.beta2d = c.beta2; aa.step = h->step_dev; aa.stats = stats;
    ppo_adam_kernel<<<nred, 256, 0, st>>>(aa);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" int dart_ppo_export_grad(dart_ppo_handle h, float* dst, void* stream) {
    if (!h || !dst) return DART_ERR_ARG;
    return cudaMemcpyAsync(dst, h->grad, NP * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream) == cudaSuccess
               ? DART_OK : DART_ERR_CUDA;
}

extern "C" int dart_ppo_apply_grad(dart_ppo_handle h, const float* src, double scale, float* stats, void* stream) {
    if (!h || !src) return DART_ERR_ARG;
    { int cur = -1; if (cudaGetDevice(&cur) != cudaSuccess || cur != h->device) return DART_ERR_ARG; }   // launch from the handle's device
    cudaStream_t st = (cudaStream_t)stream;
    const dart_ppo_cfg& c = h->cfg;
    const int nred = (NP + 255) / 256;
    ppo_grad_import_kernel<<<nred, 256, 0, st>>>(src, (float)scale, h->grad, h->normpart, h->step_dev);
    h->launches += 1;
    if (cudaGetLastError() != cudaSuccess) return DART_ERR_CUDA;
    AdamArgs aa;
    aa.nparts = nred; aa.normpart = h->normpart; aa.grad = h->grad; aa.param = h->param; aa.m = h->m; aa.v = h->v;
    aa.max_norm = (float)c.max_grad_norm; aa.wd = (float)c.weight_decay; aa.beta1 = (float)c.beta1; aa.beta2 = (float)c.beta2;
    aa.eps = (float)c.adam_eps; aa.lr = c.lr; aa.beta1d = c.beta1; aa.beta2d = c.beta2; aa.step = h->step_dev; aa.stats = stats;
    ppo_adam_kernel<<<nred, 256, 0, st>>>(aa);
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}
