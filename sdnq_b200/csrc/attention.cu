// K9: SDNQ quantized attention forward for sm_100a.
//
// What it computes is the reference's `sdnq_attn_kernel` (kernels/triton_atten.py:143-335, host side sdnq_atten_fwd :338-386):
// for every (batch z, head h) and block of query rows, a flash-attention sweep over the keys with
//     qk      = ((dot(q_q, k_q) * q_scale[m]) * k_scale[n]) * (sm_scale * log2 e)                 (:264-275)
//     mask    : causal (m >= n), boolean (int8: 0 = drop) or additive float, keys past KN             (:277-285)
//     online softmax in base 2 with running max m_i and sum l_i (l_i starts at 1, m_i at -inf)        (:231-233, 286-296)
//     acc     = acc * alpha + dot(p.to(v.dtype), v)                                                   (:297-321, unquantised PV)
//     out     = acc / l_i ; lse = m_i + log2(l_i)                                                     (:324-335)
// with 1-byte q / k codes (int8 or float8_e4m3fn) and their per-row f32 scales as `quantize_attn` (:443-487) produces them.
//
// How (nothing like the Triton program): one CTA owns 128 query rows of one head and walks the keys in tiles of 128.
//   warp 0      TMA producer: Q tile once, K tiles into a 3-deep ring, V^T tiles (two 64-key slabs) into a 2-deep ring
//   warp 1      MMA issuer:  S_j = Q K_j^T  (tcgen05 kind::i8 / kind::f8f6f4, 128x128x128, accumulator in TMEM, double-buffered)
//                            O_j = P_j V_j  (tcgen05 kind::f16, A = P_j written by the softmax warps into swizzled shared memory,
//                                            B = V^T tile; fresh accumulator per tile, double-buffered)
//   warps 2-3   idle: they only donate their registers (setmaxnreg)
//   warps 4-11  softmax: thread = one query row (its TMEM lane) x one of kParts = 2 column parts of the tile; the warps of a lane
//               quarter exchange their part-row maxima through shared memory; each keeps its own partial sum (same running maximum
//               => the partial sums simply add at the end) and its own share of the output columns in registers:
//               acc = acc * alpha + O_j is applied one tile late, while the tensor core already works on the next tile.
//               Quantised P.V (:298-318, template parameter PV): p * v_scale, a second exchange for its row maximum, 1-byte codes of
//               P written in place of the 16-bit values, O_j scaled by the tile's p_scale when it is folded.
//               Instantiations without a mask tensor / causality (kMask = false) carry no mask code in the key loop and compute a
//               quarter of their exponentials on the FMA pipe (poly_exp2) instead of the MUFU pipe.
// V is consumed K-major ([head_dim, keys]): a transposing pre-pass writes V^T into the caller's workspace once per call.
#include <mutex>

#include "act_quant.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace sdnq {
namespace {

constexpr int kBM = 128;                 // query rows per CTA
constexpr int kBN = 128;                 // keys per tile
constexpr int kKStages = 3;
constexpr int kVStages = 2;
constexpr int kTileQK = kBM * 128;       // one Q or K tile: 128 rows x 128 bytes (head dims < 128 are zero-filled by TMA)
#ifndef SDNQ_ATTN_PARTS
#define SDNQ_ATTN_PARTS 2          // 4 (16 softmax warps, 32 columns each) measured 5 % slower: profiles/r02_attention_exp_phase_experiments.md
#endif
constexpr int kParts = SDNQ_ATTN_PARTS;              // softmax threads per query row: each owns 128 / kParts of a tile's key columns
constexpr int kCPT = kBN / kParts;                   // score columns per softmax thread
constexpr int kSoftmaxThreads = 128 * kParts;
constexpr int kThreads = 128 + kSoftmaxThreads;      // warps 0-3: TMA producer, MMA issuer, two idle (register donors); then 4 * kParts softmax warps
// setmaxnreg moves registers inside the CTA's own allocation (kThreads x the launch count: 168 at 384 threads, 96 at 640): what the four
// control warps give up must cover what the softmax warps take, or the .inc waits for ever
constexpr int kLaunchRegs = kParts == 2 ? 168 : 96, kCtrlRegs = kParts == 2 ? 48 : 32, kSoftmaxRegs = kParts == 2 ? 224 : 112;
static_assert(128 * (kLaunchRegs - kCtrlRegs) >= 128 * kParts * (kSoftmaxRegs - kLaunchRegs), "setmaxnreg budget");
static_assert(kParts == 2 || kParts == 4, "column parts per row");

struct AttnParams {
    const float* q_scale;
    const float* k_scale;
    const float* v_scale;                // quantised P.V only: per-key scales of the v codes
    void* out;
    void* lse;
    const void* mask;
    int mask_kind;                       // 0 none, 1 int8 (0 = masked out), 2 f32 additive
    int64_t mask_sz, mask_sh, mask_sq, mask_sk;
    int Z, H, KH, VH, QN, KN;
    int out_dtype;
    int causal;
    float log2_scale;                    // sm_scale * log2(e)
};

// PV: 0 = 16-bit P and V (pv_matmul_dtype = None), 1 = int8 codes, 2 = e4m3 codes (per-tile row scale of P, per-key scale of V)
template <int HDV, int PV>
struct AttnCfg {
    static constexpr int kVStage = (PV == 0 ? 2 : 1) * HDV * 128;    // [HDV rows x 128 keys]: two 64-key slabs of 16-bit values, or one slab of bytes
    static constexpr int kPTile = (PV == 0 ? 2 : 1) * kBM * 128;     // P tile, same slab structure
    static constexpr int kOffK = kTileQK;
    static constexpr int kOffV = kOffK + kKStages * kTileQK;
    static constexpr int kOffP = kOffV + kVStages * kVStage;
    static constexpr int kOffTail = kOffP + 2 * kPTile;
    static constexpr int kTailBytes = 4 * kParts * kCPT * 4 * 2 + 2 * 2 * kParts * 128 * 4 + 32 * 8 + 16;
    static constexpr int kSmemBytes = kOffTail + kTailBytes + 1024;       // + alignment slack
};

enum Bar { Q_FULL = 0, K_FULL = 1, K_EMPTY = K_FULL + kKStages, V_FULL = K_EMPTY + kKStages, V_EMPTY = V_FULL + kVStages,
           S_FULL = V_EMPTY + kVStages, S_EMPTY = S_FULL + 2, P_FULL = S_EMPTY + 2, P_EMPTY = P_FULL + 2, O_FULL = P_EMPTY + 2,
           O_EMPTY = O_FULL + 2, NUM_BARS = O_EMPTY + 2 };
static_assert(NUM_BARS <= 32, "barrier area");

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#ifndef SDNQ_ATTN_POLY_PAIRS
#define SDNQ_ATTN_POLY_PAIRS 1          // of every four score pairs, how many take the FMA-pipe exponential (measured: profiles/r02_attention_exp_phase_experiments.md)
#endif
// 2^x for x <= 0 on the FMA pipe (no MUFU): n = round(x) through the 1.5 * 2^23 magic add, a degree-3 minimax polynomial of 2^f on
// f = x - n in [-0.5, 0.5] (relative error 7.5e-5, far below the 16-bit rounding of P), n added to the exponent field.  x is clamped at
// -125 so the exponent stays normal: a padded key (-inf) gets 2.4e-38 instead of 0, and meets a zero-filled V row.
__device__ __forceinline__ float poly_exp2(float x) {
    x = fmaxf(x, -125.0f);
    const float xr = x + 12582912.0f;
    const float f = x - (xr - 12582912.0f);
    float pl = fmaf(0.05517210811376572f, f, 0.2426111400127411f);
    pl = fmaf(pl, f, 0.6932608485221863f);
    pl = fmaf(pl, f, 0.9999280571937561f);
    return __int_as_float(__float_as_int(pl) + (__float_as_int(xr) << 23));
}
// the same for a pair of scores with the packed f32x2 instructions (FADD2 / FFMA2: one issue slot for both)
__device__ __forceinline__ float2 poly_exp2x2(float2 x) {
    x = make_float2(fmaxf(x.x, -125.0f), fmaxf(x.y, -125.0f));
    const float2 magic = make_float2(12582912.0f, 12582912.0f);
    const float2 xr = __fadd2_rn(x, magic);
    const float2 n = __fadd2_rn(xr, make_float2(-12582912.0f, -12582912.0f));
    const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
    float2 pl = __ffma2_rn(make_float2(0.05517210811376572f, 0.05517210811376572f), f, make_float2(0.2426111400127411f, 0.2426111400127411f));
    pl = __ffma2_rn(pl, f, make_float2(0.6932608485221863f, 0.6932608485221863f));
    pl = __ffma2_rn(pl, f, make_float2(0.9999280571937561f, 0.9999280571937561f));
    return make_float2(__int_as_float(__float_as_int(pl.x) + (__float_as_int(xr.x) << 23)), __int_as_float(__float_as_int(pl.y) + (__float_as_int(xr.y) << 23)));
}

// kMask: the call has a mask tensor or is causal; the plain instantiation carries none of that code in its key loop
template <bool kInt8, int HDV, bool kBf16, int PV, bool kMask>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_vt, const AttnParams p) {
    using C = AttnCfg<HDV, PV>;
    constexpr int kPTile = C::kPTile;
    constexpr int HALF = HDV / kParts;                     // output columns per softmax thread
    const int mask_kind = kMask ? p.mask_kind : 0;
    const bool causal = kMask && p.causal != 0;
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw = ptx::smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_dyn + (base - raw);
    const uint32_t smem_q = base, smem_k = base + C::kOffK, smem_v = base + C::kOffV, smem_p = base + C::kOffP;
    float* s_ks = reinterpret_cast<float*>(smem + C::kOffTail);          // [softmax warps][kCPT] k_scale * log2_scale of the warp's keys
    float* s_vs = s_ks + 4 * kParts * kCPT;                              // [softmax warps][kCPT] v_scale of the warp's keys (quantised P.V)
    float* s_mx = s_vs + 4 * kParts * kCPT;                              // [2][kParts][128] part-row maxima (tile parity, part, row)
    float* s_px = s_mx + 2 * kParts * 128;                               // [2][kParts][128] part-row maxima of p * v_scale (quantised P.V)
    const uint32_t bar_base = base + C::kOffTail + 4 * kParts * kCPT * 4 * 2 + 2 * 2 * kParts * 128 * 4;
    auto bar = [&](int i) { return bar_base + 8u * uint32_t(i); };
    const uint32_t tmem_slot = bar_base + 32 * 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = int(blockIdx.x) * kBM;
    const int h = blockIdx.y, z = blockIdx.z;
    const int kh = int((int64_t(h) * p.KH) / p.H), vh = int((int64_t(h) * p.VH) / p.H);     // :197-198
    const int64_t q_row0 = (int64_t(z) * p.H + h) * p.QN;                 // first row of this head in the [Z*H*QN, HD] view
    const int64_t k_row0 = (int64_t(z) * p.KH + kh) * p.KN;
    const int64_t vt_row0 = (int64_t(z) * p.VH + vh) * HDV;               // first row of this head in the [Z*VH*HDV, KN] view of V^T
    const int num_kv = (p.KN + kBN - 1) / kBN;
    const int T = causal ? min(num_kv, m0 / kBN + 1) : num_kv;          // :238-239: tiles starting past the block's last row are skipped

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_q);
        ptx::prefetch_tmap(&tmap_k);
        ptx::prefetch_tmap(&tmap_vt);
        ptx::mbar_init(bar(Q_FULL), 1);
        for (int s = 0; s < kKStages; ++s) { ptx::mbar_init(bar(K_FULL + s), 1); ptx::mbar_init(bar(K_EMPTY + s), 1); }
        for (int s = 0; s < kVStages; ++s) { ptx::mbar_init(bar(V_FULL + s), 1); ptx::mbar_init(bar(V_EMPTY + s), 1); }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(bar(S_FULL + s), 1);
            ptx::mbar_init(bar(S_EMPTY + s), kSoftmaxThreads / 32);      // one elected arrival per softmax warp
            ptx::mbar_init(bar(P_FULL + s), kSoftmaxThreads / 32);
            ptx::mbar_init(bar(P_EMPTY + s), 1);
            ptx::mbar_init(bar(O_FULL + s), 1);
            ptx::mbar_init(bar(O_EMPTY + s), kSoftmaxThreads / 32);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_launch_dependents();
    // TMEM columns: S_0 [0,128)  S_1 [128,256)  O_0 [256,256+HDV)  O_1 [384,384+HDV)
    auto s_col = [&](int b) { return uint32_t(b) * 128u; };
    auto o_col = [&](int b) { return 256u + uint32_t(b) * 128u; };

    // the four control warps need a few dozen registers, the softmax threads ~220 (64 scores + 64 output columns + staging)
    // (setmaxnreg sits at the top of each role's own branch: ptxas bounds every instruction by the smallest count that can reach it)
    if (warp < 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtrlRegs));
      if (warp == 0) {
        // ======================================================== TMA producer
        if (lane == 0) {
            pdl_wait();
            ptx::mbar_arrive_expect_tx(bar(Q_FULL), kTileQK);
            ptx::tma_load_2d(smem_q, &tmap_q, bar(Q_FULL), 0, int(q_row0 + m0));
            for (int j = 0; j < T; ++j) {
                const int ks = j % kKStages, ku = j / kKStages;
                ptx::mbar_wait_relaxed(bar(K_EMPTY + ks), (ku & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(bar(K_FULL + ks), kTileQK);
                ptx::tma_load_2d(smem_k + ks * kTileQK, &tmap_k, bar(K_FULL + ks), 0, int(k_row0 + int64_t(j) * kBN));
                const int vs = j % kVStages, vu = j / kVStages;
                ptx::mbar_wait_relaxed(bar(V_EMPTY + vs), (vu & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(bar(V_FULL + vs), C::kVStage);
                ptx::tma_load_2d(smem_v + vs * C::kVStage, &tmap_vt, bar(V_FULL + vs), j * kBN, int(vt_row0));
                if constexpr (PV == 0) ptx::tma_load_2d(smem_v + vs * C::kVStage + HDV * 128, &tmap_vt, bar(V_FULL + vs), j * kBN + 64, int(vt_row0));
            }
        }
      } else if (warp == 1) {
        // ======================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc_qk = kInt8 ? ptx::make_idesc(2, 1, 1, kBM, kBN) : ptx::make_idesc(1, 0, 0, kBM, kBN);
            const uint32_t idesc_pv = PV == 0 ? ptx::make_idesc(1, kBf16 ? 1 : 0, kBf16 ? 1 : 0, kBM, HDV)
                                      : PV == 1 ? ptx::make_idesc(2, 1, 1, kBM, HDV) : ptx::make_idesc(1, 0, 0, kBM, HDV);
            auto issue_qk = [&](int j) {
                const int ks = j % kKStages, ku = j / kKStages, b = j & 1, u = j >> 1;
                ptx::mbar_wait_relaxed(bar(K_FULL + ks), ku & 1);
                ptx::mbar_wait_relaxed(bar(S_EMPTY + b), (u & 1) ^ 1);            // the softmax warps have read S of tile j - 2
                ptx::tc_fence_after();
                const uint64_t a_desc = ptx::make_smem_desc_sw128(smem_q);
                const uint64_t b_desc = ptx::make_smem_desc_sw128(smem_k + ks * kTileQK);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    ptx::umma_ss<kInt8>(tmem_base + s_col(b), a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc_qk, k != 0 ? 1u : 0u);
                ptx::umma_commit(bar(K_EMPTY + ks));
                ptx::umma_commit(bar(S_FULL + b));
            };
            auto issue_pv = [&](int j) {
                const int vs = j % kVStages, vu = j / kVStages, b = j & 1, u = j >> 1;
                ptx::mbar_wait_relaxed(bar(V_FULL + vs), vu & 1);
                ptx::mbar_wait_relaxed(bar(P_FULL + b), u & 1);                   // P_j is in shared memory
                ptx::mbar_wait_relaxed(bar(O_EMPTY + b), (u & 1) ^ 1);            // O of tile j - 2 has been folded into the registers
                ptx::tc_fence_after();
                if constexpr (PV == 0) {
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const uint64_t a_desc = ptx::make_smem_desc_sw128(smem_p + b * kPTile + s * (kPTile / 2));
                        const uint64_t b_desc = ptx::make_smem_desc_sw128(smem_v + vs * C::kVStage + s * (HDV * 128));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ptx::umma_f16(tmem_base + o_col(b), a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc_pv, (s | k) != 0 ? 1u : 0u);
                    }
                } else {                                                  // 1-byte P and V: the 128 keys are one 128-byte slab, four K = 32 MMAs
                    const uint64_t a_desc = ptx::make_smem_desc_sw128(smem_p + b * kPTile);
                    const uint64_t b_desc = ptx::make_smem_desc_sw128(smem_v + vs * C::kVStage);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_ss<PV == 1>(tmem_base + o_col(b), a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc_pv, k != 0 ? 1u : 0u);
                }
                ptx::umma_commit(bar(V_EMPTY + vs));
                ptx::umma_commit(bar(P_EMPTY + b));
                ptx::umma_commit(bar(O_FULL + b));
            };
            ptx::mbar_wait_relaxed(bar(Q_FULL), 0);
            issue_qk(0);
            for (int j = 0; j < T; ++j) {
                if (j + 1 < T) issue_qk(j + 1);
                issue_pv(j);
            }
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kSoftmaxRegs));
        // ======================================================== softmax / output (256 threads)
        const int sw = warp - 4;                   // 0 .. 4 * kParts - 1
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int half = sw >> 2;                  // column part: which kCPT of the tile's 128 columns (and which share of the output columns)
        const int r = q * 32 + lane;               // row of the tile = TMEM lane
        const int m = m0 + r;
        const bool m_ok = m < p.QN;
        pdl_wait();
        const float qs = m_ok ? p.q_scale[q_row0 + m] : 0.f;
        // scores are kept as u = acc * k_scale' and the row's factor is applied inside the exponent: p = exp2(fma(u, rs, -m)).
        // rs = q_scale (>= 0; an all-zero row has u == 0 everywhere, so 1 serves and keeps 0 * -inf out); additive masks live in
        // the scaled domain, so with one of those the factor is applied first and rs = 1
        const float rs = mask_kind == 2 ? 1.0f : (qs > 0.f ? qs : 1.0f);
        const float pre = mask_kind == 2 ? qs : 1.0f;
        const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
        const int64_t mask_row = mask_kind != 0 ? int64_t(z) * p.mask_sz + int64_t(h) * p.mask_sh + int64_t(m_ok ? m : 0) * p.mask_sq : 0;
        float m_i = -INFINITY;
        float l_i = half == 0 ? 1.0f : 0.0f;       // :232 (l_i = 1): held by the first half, the halves add up at the end
        float alpha_pend = 0.f;
        [[maybe_unused]] float ps_pend = 1.f;      // quantised P.V: the pending tile's row scale of P
        float acc[HALF];
#pragma unroll
        for (int i = 0; i < HALF; ++i) acc[i] = 0.f;

        auto fold_o = [&](int b, int u, float alpha, [[maybe_unused]] float ps) {       // acc = acc * alpha + O_b [* p_scale]   (:297, :307, :321)
            ptx::mbar_wait(bar(O_FULL + b), u & 1);
            ptx::tc_fence_after();
            constexpr int OW = HALF < 32 ? HALF : 32;                    // columns per tcgen05.ld
#pragma unroll
            for (int c = 0; c < HALF / OW; ++c) {
                uint32_t o[OW];
                if constexpr (OW == 32) ptx::tmem_ld32(t_lane + o_col(b) + uint32_t(half * HALF + c * 32), o);
                else ptx::tmem_ld16(t_lane + o_col(b) + uint32_t(half * HALF), o);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < OW; ++i) {
                    if constexpr (PV == 0) {
                        if (i & 1) continue;                              // pairs: FFMA2
                        const float2 a2 = __ffma2_rn(make_float2(acc[c * OW + i], acc[c * OW + i + 1]), make_float2(alpha, alpha),
                                                     make_float2(__uint_as_float(o[i]), __uint_as_float(o[i + 1])));
                        acc[c * OW + i] = a2.x;
                        acc[c * OW + i + 1] = a2.y;
                    } else if constexpr (PV == 1) acc[c * OW + i] = fmaf(__int2float_rn(static_cast<int>(o[i])), ps, acc[c * OW + i] * alpha);
                    else acc[c * OW + i] = fmaf(__uint_as_float(o[i]), ps, acc[c * OW + i] * alpha);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar(O_EMPTY + b));
        };

        // the warp's 64 key scales per tile travel global -> registers (one tile ahead) -> the warp's own strip of shared memory: a load
        // issued in the tile it is needed in would put an L2 round trip in front of every tile, a CTA-wide staging barrier would
        // keep all eight warps in lock-step (they then fight for the MUFU pipe at the same moment)
        float* my_ks = s_ks + sw * kCPT;
        auto load_ks = [&](int j, int e) { const int n = j * kBN + half * kCPT + e * 32 + lane; return (e * 32 < kCPT && n < p.KN) ? p.k_scale[k_row0 + n] : 0.f; };
        float ks_a = load_ks(0, 0), ks_b = load_ks(0, 1);                 // (ks_b / vs_b: the second 32 keys of a 64-column part, unused with 32)
        [[maybe_unused]] float* my_vs = s_vs + sw * kCPT;
        const int64_t v_row0 = (int64_t(z) * p.VH + vh) * p.KN;
        auto load_vs = [&](int j, int e) { const int n = j * kBN + half * kCPT + e * 32 + lane; return (PV != 0 && e * 32 < kCPT && n < p.KN) ? p.v_scale[v_row0 + n] : 0.f; };
        [[maybe_unused]] float vs_a = load_vs(0, 0), vs_b = load_vs(0, 1);
        for (int j = 0; j < T; ++j) {
            const int b = j & 1, u = j >> 1;
            const int n0 = j * kBN;
            my_ks[lane] = ks_a * p.log2_scale;
            if constexpr (kCPT == 64) my_ks[32 + lane] = ks_b * p.log2_scale;
            if constexpr (PV != 0) { my_vs[lane] = vs_a; if constexpr (kCPT == 64) my_vs[32 + lane] = vs_b; }
            __syncwarp();
            if (j + 1 < T) {
                ks_a = load_ks(j + 1, 0); ks_b = load_ks(j + 1, 1);
                if constexpr (PV != 0) { vs_a = load_vs(j + 1, 0); vs_b = load_vs(j + 1, 1); }
            }
            ptx::mbar_wait(bar(S_FULL + b), u & 1);
            ptx::tc_fence_after();
            float t[kCPT];
            {
                uint32_t s0[32];
                [[maybe_unused]] uint32_t s1[32];
                ptx::tmem_ld32(t_lane + s_col(b) + uint32_t(half * kCPT), s0);
                if constexpr (kCPT == 64) ptx::tmem_ld32(t_lane + s_col(b) + uint32_t(half * kCPT + 32), s1);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar(S_EMPTY + b));
                const float4* ks4 = reinterpret_cast<const float4*>(my_ks);
#pragma unroll
                for (int i4 = 0; i4 < kCPT / 4; ++i4) {
                    const float4 kv = ks4[i4];
                    const float kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
                    for (int e = 0; e < 4; e += 2) {                      // two scores per FMUL2 (packed f32x2: one issue slot)
                        const int i = i4 * 4 + e;
                        const uint32_t rv0 = i < 32 ? s0[i & 31] : s1[i & 31], rv1 = i < 32 ? s0[(i + 1) & 31] : s1[(i + 1) & 31];
                        float2 a;
                        if constexpr (kInt8) a = make_float2(__int2float_rn(static_cast<int>(rv0)), __int2float_rn(static_cast<int>(rv1)));      // I2FP: one issue slot (the magic-number add takes two)
                        else a = make_float2(__uint_as_float(rv0), __uint_as_float(rv1));
                        const float2 tt = __fmul2_rn(a, make_float2(kk[e], kk[e + 1]));
                        t[i] = tt.x;
                        t[i + 1] = tt.y;
                    }
                }
                __syncwarp();                                             // my_ks is rewritten at the top of the next tile
            }
            // masks
            const int col0 = n0 + half * kCPT;
            if (mask_kind == 1) {
                const int8_t* mk = reinterpret_cast<const int8_t*>(p.mask) + mask_row;
#pragma unroll
                for (int i = 0; i < kCPT; ++i) {
                    const int n = col0 + i;
                    if (n < p.KN && mk[int64_t(n) * p.mask_sk] == 0) t[i] = -INFINITY;
                }
            } else if (mask_kind == 2) {
                const float* mk = reinterpret_cast<const float*>(p.mask) + mask_row;
#pragma unroll
                for (int i = 0; i < kCPT; ++i) {
                    const int n = col0 + i;
                    t[i] *= pre;
                    if (n < p.KN) t[i] += mk[int64_t(n) * p.mask_sk];
                }
            }
            if (col0 + kCPT > p.KN || (causal && col0 + kCPT - 1 > m)) {
#pragma unroll
                for (int i = 0; i < kCPT; ++i) {
                    const int n = col0 + i;
                    if (n >= p.KN || (causal && n > m)) t[i] = -INFINITY;
                }
            }
            float mx4[4] = {t[0], t[1], t[2], t[3]};
#pragma unroll
            for (int i = 4; i < kCPT; i += 4) {
                mx4[0] = fmaxf(mx4[0], t[i]); mx4[1] = fmaxf(mx4[1], t[i + 1]); mx4[2] = fmaxf(mx4[2], t[i + 2]); mx4[3] = fmaxf(mx4[3], t[i + 3]);
            }
            float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            s_mx[(b * kParts + half) * 128 + r] = mx;
            asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * kParts) : "memory");
#pragma unroll
            for (int o = 1; o < kParts; ++o) mx = fmaxf(mx, s_mx[(b * kParts + ((half + o) & (kParts - 1))) * 128 + r]);
            const float m_new = fmaxf(m_i, mx * rs);
            // :287-294 -- one formula for both of the reference's branches: exp2(-inf - finite) = 0, and a row that has seen
            // nothing but masked keys keeps alpha = 1, p = 0
            const float alpha = m_new == -INFINITY ? 1.0f : fast_exp2(m_i - m_new);
            const float m_use = m_new == -INFINITY ? 0.0f : m_new;
            float sum4[4] = {0.f, 0.f, 0.f, 0.f};
            [[maybe_unused]] float2 sum2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};      // 16-bit P.V: pair sums (FADD2)
            ptx::mbar_wait(bar(P_EMPTY + b), (u & 1) ^ 1);               // the MMAs of tile j - 2 have read this P buffer
            [[maybe_unused]] float ps = 1.f;
            if constexpr (PV == 0) {
                // the thread's kCPT columns inside the two 64-key slabs of the tile: slab (half * kCPT) / 64, first 16-byte chunk ((half * kCPT) % 64) / 8
                const uint32_t p_row = smem_p + b * kPTile + uint32_t((half * kCPT) / 64) * (kPTile / 2) + uint32_t(r) * 128u;
                const uint32_t cb = uint32_t(((half * kCPT) % 64) / 8);
#pragma unroll
                for (int c = 0; c < kCPT / 8; ++c) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        // the exponentials are MUFU-bound (16 / clock / SM) with issue slots to spare: without a mask (no -inf scores
                        // except padded keys, which meet zero-filled V rows) a share of them is computed on the FMA pipe instead
                        const bool kPoly = !kMask && e >= 4 - SDNQ_ATTN_POLY_PAIRS;      // (a compile-time constant once the loop is unrolled)
                        const float2 x = __ffma2_rn(make_float2(t[c * 8 + 2 * e], t[c * 8 + 2 * e + 1]), make_float2(rs, rs), make_float2(-m_use, -m_use));
                        float p0, p1;
                        if (kPoly) {
                            const float2 pp = poly_exp2x2(x);
                            p0 = pp.x;
                            p1 = pp.y;
                        } else {
                            p0 = fast_exp2(x.x);
                            p1 = fast_exp2(x.y);
                        }
                        sum2[e & 1] = __fadd2_rn(sum2[e & 1], make_float2(p0, p1));
                        if constexpr (kBf16) {
                            __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
                            w[e] = *reinterpret_cast<uint32_t*>(&hh);
                        } else {
                            __half2 hh = __floats2half2_rn(p0, p1);
                            w[e] = *reinterpret_cast<uint32_t*>(&hh);
                        }
                    }
                    ptx::st_shared_v4(p_row + (((cb + uint32_t(c)) ^ uint32_t(r & 7)) << 4), w[0], w[1], w[2], w[3]);
                }
            } else {
                // :298-318: p *= v_scale; p_scale = rowmax(p) / 127 (or / 448), 1 if it underflows; codes = floor(p / p_scale + 0.5)
                // (int8) or the e4m3 rounding of p / p_scale.  The row maximum spans both halves of the tile: a second exchange.
                const float4* vs4 = reinterpret_cast<const float4*>(my_vs);
                float px4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i4 = 0; i4 < kCPT / 4; ++i4) {
                    const float4 vv = vs4[i4];
                    const float vk[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float pe = fast_exp2(fmaf(t[i4 * 4 + e], rs, -m_use));
                        sum4[e] += pe;
                        t[i4 * 4 + e] = pe * vk[e];
                        px4[e] = fmaxf(px4[e], t[i4 * 4 + e]);
                    }
                }
                float px = fmaxf(fmaxf(px4[0], px4[1]), fmaxf(px4[2], px4[3]));
                s_px[(b * kParts + half) * 128 + r] = px;
                asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * kParts) : "memory");
#pragma unroll
                for (int o = 1; o < kParts; ++o) px = fmaxf(px, s_px[(b * kParts + ((half + o) & (kParts - 1))) * 128 + r]);
                ps = px * (PV == 1 ? (1.0f / 127.0f) : (1.0f / 448.0f));
                if (ps <= 2e-38f) ps = 1.0f;
                const float inv = __fdiv_rn(1.0f, ps);
                const uint32_t p_row = smem_p + b * kPTile + uint32_t(r) * 128u;
#pragma unroll
                for (int c = 0; c < kCPT / 16; ++c) {                     // 16 codes per 16-byte chunk
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float* tv = &t[c * 16 + e * 4];
                        if constexpr (PV == 1) {
                            // floor(p / p_scale + 0.5) in [0, 127] without F2I (which shares the MUFU pipe's rate with the exponentials):
                            // a round-down add of 2^23 leaves the integer in the low mantissa bits, PRMT gathers the four low bytes
                            const uint32_t b0 = __float_as_uint(__fadd_rd(fmaf(tv[0], inv, 0.5f), 8388608.0f)), b1 = __float_as_uint(__fadd_rd(fmaf(tv[1], inv, 0.5f), 8388608.0f));
                            const uint32_t b2 = __float_as_uint(__fadd_rd(fmaf(tv[2], inv, 0.5f), 8388608.0f)), b3 = __float_as_uint(__fadd_rd(fmaf(tv[3], inv, 0.5f), 8388608.0f));
                            w[e] = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
                        } else {
                            const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(tv[0] * inv, tv[1] * inv), __NV_SATFINITE, __NV_E4M3);
                            const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(tv[2] * inv, tv[3] * inv), __NV_SATFINITE, __NV_E4M3);
                            w[e] = lo | (hi << 16);
                        }
                    }
                    ptx::st_shared_v4(p_row + (uint32_t((half * (kCPT / 16) + c) ^ (r & 7)) << 4), w[0], w[1], w[2], w[3]);
                }
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar(P_FULL + b));
            if constexpr (PV == 0) l_i = fmaf(l_i, alpha, (sum2[0].x + sum2[0].y) + (sum2[1].x + sum2[1].y));
            else l_i = fmaf(l_i, alpha, (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
            m_i = m_new;
            if (j > 0) fold_o(b ^ 1, (j - 1) >> 1, alpha_pend, ps_pend);
            alpha_pend = alpha;
            ps_pend = ps;
        }
        fold_o((T - 1) & 1, (T - 1) >> 1, alpha_pend, ps_pend);
        // total row sum = the two halves' partial sums (same running maximum)
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * kParts) : "memory");        // the last tile's s_mx reads are done
        s_mx[half * 128 + r] = l_i;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * kParts) : "memory");
        float l_tot = 0.f;
#pragma unroll
        for (int o = 0; o < kParts; ++o) l_tot += s_mx[o * 128 + r];      // the same order in every part: all of a row's threads divide by the same sum
        if (m_ok) {
            const float inv = 1.0f / l_tot;                                // :324
            const int64_t orow = (q_row0 + m) * HDV + half * HALF;
            if (p.out_dtype == SDNQ_F32) {
                float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow);
#pragma unroll
                for (int i = 0; i < HALF / 4; ++i) dst[i] = make_float4(acc[4 * i] * inv, acc[4 * i + 1] * inv, acc[4 * i + 2] * inv, acc[4 * i + 3] * inv);
            } else {
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + orow);
#pragma unroll
                for (int i = 0; i < HALF / 8; ++i) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float a0 = acc[8 * i + 2 * e] * inv, a1 = acc[8 * i + 2 * e + 1] * inv;
                        if (p.out_dtype == SDNQ_BF16) {
                            __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
                            w[e] = *reinterpret_cast<uint32_t*>(&hh);
                        } else {
                            __half2 hh = __floats2half2_rn(a0, a1);
                            w[e] = *reinterpret_cast<uint32_t*>(&hh);
                        }
                    }
                    dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            if (p.lse != nullptr && half == 0) {                           // :326-332
                float l = m_i + log2f(l_tot);
                if (mask_kind != 0 && l == -INFINITY) l = 0.f;
                const int64_t li = q_row0 + m;
                if (p.out_dtype == SDNQ_F32) reinterpret_cast<float*>(p.lse)[li] = l;
                else if (p.out_dtype == SDNQ_BF16) reinterpret_cast<__nv_bfloat16*>(p.lse)[li] = __float2bfloat16_rn(l);
                else reinterpret_cast<__half*>(p.lse)[li] = __float2half_rn(l);
            }
        }
    }
    // ---- teardown
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// V [R, N, HD] (16-bit values or 1-byte codes) -> V^T [R, HD, ldt] (row pitch ldt >= N, 16-byte multiple); columns N..ldt are never
// read (the tensor map ends at N)
template <typename E>
__global__ void __launch_bounds__(256) transpose_v_kernel(const E* __restrict__ v, E* __restrict__ vt, int N, int HD, int64_t ldt) {
    __shared__ E tile[32][32 + 4 / sizeof(E)];
    pdl_launch_dependents();
    pdl_wait();
    const int64_t rr = blockIdx.z;
    const int n0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int n = n0 + i, d = d0 + tx;
        tile[i][tx] = (n < N && d < HD) ? v[(rr * N + n) * HD + d] : E(0);
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int d = d0 + i, n = n0 + tx;
        if (d < HD && n < N) vt[(rr * HD + d) * ldt + n] = tile[tx][i];
    }
}

// The 16-bit case in 64 x 64 tiles: a lane moves two values at a time (4-byte loads along the head dim, 4-byte stores along the keys),
// so both sides of the transposition touch full 128-byte lines; HD % 64 == 0, v 4-byte aligned.
__global__ void __launch_bounds__(256) transpose_v16_kernel(const uint16_t* __restrict__ v, uint16_t* __restrict__ vt, int N, int HD, int64_t ldt) {
    __shared__ uint16_t tile[64][65];                // odd pitch: the column reads of the store phase fall on 32 different banks
    pdl_launch_dependents();
    pdl_wait();
    const int64_t rr = blockIdx.z;
    const int n0 = blockIdx.x * 64, d0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = ty; i < 64; i += 8) {
        const int n = n0 + i;
        const uint32_t w = n < N ? *reinterpret_cast<const uint32_t*>(v + (rr * N + n) * HD + d0 + 2 * tx) : 0u;
        tile[i][2 * tx] = static_cast<uint16_t>(w);
        tile[i][2 * tx + 1] = static_cast<uint16_t>(w >> 16);
    }
    __syncthreads();
    const int n = n0 + 2 * tx;
    if (n < N) {
#pragma unroll
        for (int i = ty; i < 64; i += 8) {
            const uint32_t w = uint32_t(tile[2 * tx][i]) | (uint32_t(tile[2 * tx + 1][i]) << 16);
            *reinterpret_cast<uint32_t*>(vt + (rr * HD + d0 + i) * ldt + n) = w;      // (an odd N: the pair's second half is 0 and lands in the row padding)
        }
    }
}

// smooth-K (triton_atten.py:456-461): k [R, N, HD] -> k.to(f32) - mean over the N tokens, written as f32 or in a 16-bit type.
// A cluster of kMeanSplit CTAs per (batch, head): each sums a contiguous eighth of the tokens, every CTA adds the eight partial rows
// through distributed shared memory in rank order (so all hold the same mean), then subtracts it from its own tokens (second read from L2).
constexpr int kMeanSplit = 8;
template <typename TIn, typename TOut>
__global__ void __cluster_dims__(1, kMeanSplit, 1) __launch_bounds__(256) smooth_k_kernel(const TIn* __restrict__ k, TOut* __restrict__ out, int N, int HD) {
    __shared__ float s_part[256];
    __shared__ float s_tot[256];
    __shared__ float s_mean[256];
    pdl_launch_dependents();
    pdl_wait();
    const int64_t head = int64_t(blockIdx.x) * N * HD;
    const int tid = threadIdx.x;
    const int lanes = HD < 256 ? HD : 256;           // threads along the channel axis (HD <= 256)
    const int groups = 256 / lanes;                  // row groups
    const int d = tid % lanes, g = tid / lanes;
    const int chunk = (N + kMeanSplit - 1) / kMeanSplit;
    const int n_begin = int(blockIdx.y) * chunk, n_end = min(N, n_begin + chunk);
    float sum = 0.f;
    if (g < groups)
        for (int n = n_begin + g; n < n_end; n += groups) sum += ElemTraits<TIn>::load(k[head + int64_t(n) * HD + d]);
    s_part[tid] = g < groups ? sum : 0.f;
    __syncthreads();
    if (tid < lanes) {
        float tot = 0.f;
        for (int gg = 0; gg < groups; ++gg) tot += s_part[gg * lanes + tid];
        s_tot[tid] = tot;
    }
    ptx::cluster_sync();
    if (tid < lanes) {
        float tot = 0.f;
#pragma unroll
        for (int rk = 0; rk < kMeanSplit; ++rk) {
            float part;
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(part) : "r"(ptx::mapa(ptx::smem_u32(&s_tot[tid]), uint32_t(rk))));
            tot += part;
        }
        s_mean[tid] = tot / static_cast<float>(N);
    }
    ptx::cluster_sync();                             // (also: no CTA leaves while a peer may still read its partial row)
    if (g < groups) {
        const float mean = s_mean[d];
        for (int n = n_begin + g; n < n_end; n += groups) {
            const int64_t i = head + int64_t(n) * HD + d;
            const float val = ElemTraits<TIn>::load(k[i]) - mean;
            if constexpr (sizeof(TOut) == 4) out[i] = val;
            else if constexpr (std::is_same<TOut, __nv_bfloat16>::value) out[i] = __float2bfloat16_rn(val);
            else out[i] = __float2half_rn(val);
        }
    }
}

// Per-head channel means for smooth-K: k [heads, N, HD] -> mean [heads, HD] f32 (sum in f32, / N).  A cluster of kMeanSplit CTAs per
// head: each CTA sums a contiguous eighth of the tokens (a lane owns 8 consecutive channels, four independent 16-byte loads in flight
// per thread, the 256 / (HD / 8) row groups of the CTA reduced through shared memory), CTA 0 of the cluster adds the eight partial rows
// through distributed shared memory in rank order -- deterministic, no workspace, and 8 x heads CTAs instead of heads (one CTA per
// head with one load in flight ran at 1.0 TB/s: 112 us of a 1.6 ms FLUX call).
template <typename T>
__global__ void __cluster_dims__(1, kMeanSplit, 1) __launch_bounds__(256)
attn_colmean_kernel(const T* __restrict__ k, float* __restrict__ mean, int N, int HD) {
    __shared__ float s_part[256 * 8];
    __shared__ float s_tot[256];
    pdl_launch_dependents();
    pdl_wait();
    const int lpr = HD / 8, groups = 256 / lpr;
    const int sub = threadIdx.x % lpr, g = threadIdx.x / lpr;
    const int chunk = (N + kMeanSplit - 1) / kMeanSplit;
    const int n_begin = int(blockIdx.y) * chunk, n_end = min(N, n_begin + chunk);
    const T* src = k + int64_t(blockIdx.x) * N * HD + sub * 8;
    float sum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sum[i] = 0.f;
    int n = n_begin + g;
    for (; n + 3 * groups < n_end; n += 4 * groups) {
        actq::Held<T> h0, h1, h2, h3;
        h0.load(src + int64_t(n) * HD);
        h1.load(src + int64_t(n + groups) * HD);
        h2.load(src + int64_t(n + 2 * groups) * HD);
        h3.load(src + int64_t(n + 3 * groups) * HD);
        float v0[8], v1[8], v2[8], v3[8];
        h0.get(v0); h1.get(v1); h2.get(v2); h3.get(v3);
#pragma unroll
        for (int i = 0; i < 8; ++i) sum[i] += (v0[i] + v1[i]) + (v2[i] + v3[i]);
    }
    for (; n < n_end; n += groups) {
        actq::Held<T> hv;
        hv.load(src + int64_t(n) * HD);
        float v[8];
        hv.get(v);
#pragma unroll
        for (int i = 0; i < 8; ++i) sum[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_part[threadIdx.x * 8 + i] = sum[i];
    __syncthreads();
    if (int(threadIdx.x) < HD) {
        const int c = threadIdx.x, csub = c / 8, ci = c % 8;
        float tot = 0.f;
        for (int gg = 0; gg < groups; ++gg) tot += s_part[(gg * lpr + csub) * 8 + ci];
        s_tot[c] = tot;
    }
    ptx::cluster_sync();                                        // every CTA's partial row is in its shared memory
    if (blockIdx.y == 0 && int(threadIdx.x) < HD) {
        float tot = 0.f;
#pragma unroll
        for (int rk = 0; rk < kMeanSplit; ++rk) {
            float part;
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(part) : "r"(ptx::mapa(ptx::smem_u32(&s_tot[threadIdx.x]), uint32_t(rk))));
            tot += part;
        }
        mean[int64_t(blockIdx.x) * HD + threadIdx.x] = tot / static_cast<float>(N);
    }
    ptx::cluster_sync();                                        // no CTA leaves while CTA 0 may still read its shared memory
}

// Row quantiser for attention operands (quantize_attn without a rotation, triton_atten.py:456-471): rows of HD = 8 * LPR values,
// LPR lanes per row (32 / LPR rows per warp and step); optional smooth-K: the head's channel means are subtracted in f32 first.
//   scale = amax / 127 (int8) or / 448 (e4m3); codes = the same exact-division quantiser as the Linear pre-pass (actq::quantise8)
template <typename T, int MODE>
__global__ void __launch_bounds__(256) attn_rowquant_kernel(const T* __restrict__ x, const float* __restrict__ mean, int rows_per_head,
                                                            int64_t rows, int HD, uint8_t* __restrict__ xq, float* __restrict__ scale) {
    pdl_launch_dependents();
    pdl_wait();
    const int lpr = HD / 8, rpw = 32 / lpr;
    const int lane = threadIdx.x & 31, sub = lane % lpr;
    const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t warps_total = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t r0 = warp_global * rpw; r0 < rows; r0 += warps_total * rpw) {
        const int64_t row = r0 + lane / lpr;
        const bool ok = row < rows;
        float v[8];
        if (ok) {
            actq::Held<T> hv;
            hv.load(x + row * HD + sub * 8);
            hv.get(v);
            if (mean != nullptr) {
                const float* mp = mean + (row / rows_per_head) * HD + sub * 8;
                const float4 m0 = *reinterpret_cast<const float4*>(mp), m1 = *reinterpret_cast<const float4*>(mp + 4);
                v[0] -= m0.x; v[1] -= m0.y; v[2] -= m0.z; v[3] -= m0.w;
                v[4] -= m1.x; v[5] -= m1.y; v[6] -= m1.z; v[7] -= m1.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) amax = fmaxf(amax, fabsf(v[i]));
        for (int o = 1; o < lpr; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        const float sc = __fdiv_rn(amax, MODE == SDNQ_F8E4M3 ? 448.f : 127.f);
        const actq::RowDivider divider(sc);
        int unused = 0;
        const uint2 codes = divider.safe() ? actq::quantise8<MODE, true>(v, divider, 0.f, false, unused)
                                           : actq::quantise8<MODE, false>(v, divider, 0.f, false, unused);
        if (ok) {
            *reinterpret_cast<uint2*>(xq + row * HD + sub * 8) = codes;
            if (sub == 0) scale[row] = sc;
        }
    }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// [rows, cols] matrix of elem_bytes-wide elements, row pitch pitch_bytes, box {128 bytes, box_rows}, 128-byte swizzle
int make_tmap_sw128(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t pitch_bytes, int elem_bytes, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    SDNQ_REQUIRE(enc != nullptr, SDNQ_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch_bytes)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elem_bytes), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(ptr), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SDNQ_REQUIRE(r == CUDA_SUCCESS, SDNQ_ECUDA, "cuTensorMapEncodeTiled (attention) failed with CUresult %d (rows=%lld cols=%lld pitch=%lld)",
                 static_cast<int>(r), (long long)rows, (long long)cols, (long long)pitch_bytes);
    return SDNQ_OK;
}

template <bool kInt8, int HDV, bool kBf16, int PV, bool kMask>
int launch_attn_m(const void* q, const void* k, const void* vt, int64_t ldt, int HD, const AttnParams& p, cudaStream_t st) {
    using C = AttnCfg<HDV, PV>;
    auto kernel = attn_fwd_kernel<kInt8, HDV, kBf16, PV, kMask>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    static int launch_regs = 0;
    std::call_once(once, [&] {
        attr_err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        cudaFuncAttributes fa{};
        if (attr_err == cudaSuccess) attr_err = cudaFuncGetAttributes(&fa, kernel);
        launch_regs = fa.numRegs;
    });
    // the register hand-over inside the kernel is sized for this count: a different one (another compiler) could leave setmaxnreg.inc waiting
    SDNQ_REQUIRE(attr_err != cudaSuccess || launch_regs == kLaunchRegs, SDNQ_ECUDA, "attn_fwd_kernel was compiled to %d registers per thread, the setmaxnreg budget assumes %d",
                 launch_regs, kLaunchRegs);
    SDNQ_REQUIRE(attr_err == cudaSuccess, SDNQ_ECUDA, "cudaFuncSetAttribute(max dynamic smem %d) failed: %s", C::kSmemBytes, cudaGetErrorString(attr_err));
    CUtensorMap tq, tk, tv;
    int rc = make_tmap_sw128(&tq, q, int64_t(p.Z) * p.H * p.QN, HD, HD, 1, kBM);
    if (rc != SDNQ_OK) return rc;
    rc = make_tmap_sw128(&tk, k, int64_t(p.Z) * p.KH * p.KN, HD, HD, 1, kBN);
    if (rc != SDNQ_OK) return rc;
    rc = make_tmap_sw128(&tv, vt, int64_t(p.Z) * p.VH * HDV, p.KN, ldt * (PV == 0 ? 2 : 1), PV == 0 ? 2 : 1, HDV);
    if (rc != SDNQ_OK) return rc;
    const dim3 grid((p.QN + kBM - 1) / kBM, p.H, p.Z);
    SDNQ_CUDA_OK(launch_pdl(kernel, grid, dim3(kThreads), size_t(C::kSmemBytes), st, tq, tk, tv, p));
    return check_launch("attn_fwd_kernel");
}

template <bool kInt8, int HDV, bool kBf16, int PV>
int launch_attn(const void* q, const void* k, const void* vt, int64_t ldt, int HD, const AttnParams& p, cudaStream_t st) {
    return (p.mask_kind != 0 || p.causal != 0) ? launch_attn_m<kInt8, HDV, kBf16, PV, true>(q, k, vt, ldt, HD, p, st)
                                               : launch_attn_m<kInt8, HDV, kBf16, PV, false>(q, k, vt, ldt, HD, p, st);
}

int64_t vt_pitch(int64_t KN, int elem_bytes) { const int64_t per = 16 / elem_bytes; return (KN + per - 1) / per * per; }

}  // namespace
}  // namespace sdnq

using namespace sdnq;

extern "C" size_t sdnq_b200_attention_workspace_bytes(int64_t Z, int64_t VH, int64_t KN, int64_t HDV) {
    if (Z <= 0 || VH <= 0 || KN <= 0 || HDV <= 0) return 0;
    return static_cast<size_t>(Z * VH * HDV * vt_pitch(KN, 2) * 2);       // (covers the 1-byte layout too)
}

extern "C" int sdnq_b200_smooth_k(const void* k, int k_dtype, int64_t heads, int64_t N, int64_t HD, void* out, int out_dtype, void* stream) {
    SDNQ_REQUIRE(k && out, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(heads >= 0 && N > 0 && HD > 0 && HD <= 256 && N < (1LL << 31) && heads < (1LL << 31), SDNQ_EUNSUPPORTED, "smooth_k: heads=%lld N=%lld HD=%lld (HD <= 256)",
                 (long long)heads, (long long)N, (long long)HD);
    if (heads == 0) return SDNQ_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const dim3 grid(static_cast<unsigned>(heads), kMeanSplit), block(256);
#define SDNQ_SMOOTH(TI, TO)                                                                                                   \
    do {                                                                                                                      \
        SDNQ_CUDA_OK(launch_pdl(smooth_k_kernel<TI, TO>, grid, block, 0, st, reinterpret_cast<const TI*>(k), reinterpret_cast<TO*>(out), int(N), int(HD))); \
        return check_launch("smooth_k_kernel");                                                                               \
    } while (0)
    if (k_dtype == SDNQ_BF16 && out_dtype == SDNQ_F32) SDNQ_SMOOTH(__nv_bfloat16, float);
    if (k_dtype == SDNQ_BF16 && out_dtype == SDNQ_BF16) SDNQ_SMOOTH(__nv_bfloat16, __nv_bfloat16);
    if (k_dtype == SDNQ_F16 && out_dtype == SDNQ_F32) SDNQ_SMOOTH(__half, float);
    if (k_dtype == SDNQ_F16 && out_dtype == SDNQ_F16) SDNQ_SMOOTH(__half, __half);
    if (k_dtype == SDNQ_F32 && out_dtype == SDNQ_F32) SDNQ_SMOOTH(float, float);
    if (k_dtype == SDNQ_F32 && out_dtype == SDNQ_BF16) SDNQ_SMOOTH(float, __nv_bfloat16);
    if (k_dtype == SDNQ_F32 && out_dtype == SDNQ_F16) SDNQ_SMOOTH(float, __half);
#undef SDNQ_SMOOTH
    return set_error(SDNQ_EUNSUPPORTED, "smooth_k: dtype pair %d -> %d", k_dtype, out_dtype);
}

extern "C" int sdnq_b200_attention(const void* q, const void* k, const void* v, int qk_dtype, int v_dtype, const float* q_scale,
                                   const float* k_scale, const float* v_scale, const void* mask, int mask_dtype, const int64_t* mask_strides, void* out,
                                   void* lse, int out_dtype, int64_t Z, int64_t H, int64_t KH, int64_t VH, int64_t QN, int64_t KN,
                                   int64_t HD, int64_t HDV, float sm_scale, int is_causal, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    SDNQ_REQUIRE(q && k && v && q_scale && k_scale && out, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(qk_dtype == SDNQ_I8 || qk_dtype == SDNQ_F8E4M3, SDNQ_EUNSUPPORTED, "attention: q / k codes must be int8 or float8_e4m3fn (got %d)", qk_dtype);
    SDNQ_REQUIRE(v_dtype == SDNQ_BF16 || v_dtype == SDNQ_F16 || v_dtype == SDNQ_I8 || v_dtype == SDNQ_F8E4M3, SDNQ_EUNSUPPORTED,
                 "attention: v must be bf16 / f16 values or int8 / float8_e4m3fn codes (got %d)", v_dtype);
    const int pv = v_dtype == SDNQ_I8 ? 1 : v_dtype == SDNQ_F8E4M3 ? 2 : 0;
    SDNQ_REQUIRE((pv != 0) == (v_scale != nullptr), SDNQ_EINVAL, "attention: v_scale goes with 1-byte v codes, and only with them");
    SDNQ_REQUIRE(out_dtype == SDNQ_BF16 || out_dtype == SDNQ_F16 || out_dtype == SDNQ_F32, SDNQ_EINVAL, "attention: bad out dtype %d", out_dtype);
    SDNQ_REQUIRE(Z > 0 && H > 0 && KH > 0 && VH > 0 && QN > 0 && KN > 0, SDNQ_EINVAL, "attention: bad shape");
    SDNQ_REQUIRE(HD % 16 == 0 && HD >= 16 && HD <= 128, SDNQ_EUNSUPPORTED, "attention: head dim of q / k must be a multiple of 16 up to 128 (got %lld)", (long long)HD);
    SDNQ_REQUIRE(HDV == 64 || HDV == 128, SDNQ_EUNSUPPORTED, "attention: head dim of v must be 64 or 128 (got %lld; pad as get_attn_inputs does)", (long long)HDV);
    SDNQ_REQUIRE(Z * H * QN < (1LL << 31) && Z * KH * KN < (1LL << 31) && Z * VH * HDV < (1LL << 31) && Z < 65536 && H < 65536 && Z * VH < 65536, SDNQ_EUNSUPPORTED,
                 "attention: too many rows for 32-bit TMA coordinates");
    SDNQ_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out) |
                   reinterpret_cast<uintptr_t>(workspace)) & 15) == 0, SDNQ_EINVAL, "attention: q, k, v, out and workspace must be 16-byte aligned");
    SDNQ_REQUIRE(workspace && workspace_bytes >= sdnq_b200_attention_workspace_bytes(Z, VH, KN, HDV), SDNQ_EINVAL, "attention: workspace too small");
    int mask_kind = 0;
    if (mask != nullptr) {
        SDNQ_REQUIRE(mask_strides != nullptr, SDNQ_EINVAL, "attention: mask without strides");
        SDNQ_REQUIRE(mask_dtype == SDNQ_I8 || mask_dtype == SDNQ_F32, SDNQ_EUNSUPPORTED, "attention: mask must be int8 (boolean) or f32 (additive), got %d", mask_dtype);
        mask_kind = mask_dtype == SDNQ_I8 ? 1 : 2;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t ldt = vt_pitch(KN, pv == 0 ? 2 : 1);
    {
        const dim3 grid(static_cast<unsigned>((KN + 31) / 32), static_cast<unsigned>((HDV + 31) / 32), static_cast<unsigned>(Z * VH));
        if (pv == 0 && HDV % 64 == 0 && (reinterpret_cast<uintptr_t>(v) & 3) == 0) {
            const dim3 grid64(static_cast<unsigned>((KN + 63) / 64), static_cast<unsigned>(HDV / 64), static_cast<unsigned>(Z * VH));
            SDNQ_CUDA_OK(launch_pdl(transpose_v16_kernel, grid64, dim3(256), 0, st, reinterpret_cast<const uint16_t*>(v), reinterpret_cast<uint16_t*>(workspace),
                                    int(KN), int(HDV), ldt));
        } else if (pv == 0)
            SDNQ_CUDA_OK(launch_pdl(transpose_v_kernel<uint16_t>, grid, dim3(256), 0, st, reinterpret_cast<const uint16_t*>(v), reinterpret_cast<uint16_t*>(workspace),
                                    int(KN), int(HDV), ldt));
        else
            SDNQ_CUDA_OK(launch_pdl(transpose_v_kernel<uint8_t>, grid, dim3(256), 0, st, reinterpret_cast<const uint8_t*>(v), reinterpret_cast<uint8_t*>(workspace),
                                    int(KN), int(HDV), ldt));
        int rc = check_launch("transpose_v_kernel");
        if (rc != SDNQ_OK) return rc;
    }
    AttnParams p{};
    p.q_scale = q_scale;
    p.k_scale = k_scale;
    p.v_scale = v_scale;
    p.out = out;
    p.lse = lse;
    p.mask = mask;
    p.mask_kind = mask_kind;
    if (mask_kind != 0) { p.mask_sz = mask_strides[0]; p.mask_sh = mask_strides[1]; p.mask_sq = mask_strides[2]; p.mask_sk = mask_strides[3]; }
    p.Z = int(Z); p.H = int(H); p.KH = int(KH); p.VH = int(VH); p.QN = int(QN); p.KN = int(KN);
    p.out_dtype = out_dtype;
    p.causal = is_causal ? 1 : 0;
    p.log2_scale = sm_scale * 1.4426950408889634f;
    const bool i8 = qk_dtype == SDNQ_I8, bf = v_dtype == SDNQ_BF16;
#define SDNQ_ATTN_PV(HDV_, BF_, PV_)                                                                                           \
    (i8 ? launch_attn<true, HDV_, BF_, PV_>(q, k, workspace, ldt, int(HD), p, st) : launch_attn<false, HDV_, BF_, PV_>(q, k, workspace, ldt, int(HD), p, st))
#define SDNQ_ATTN(HDV_)                                                                                                        \
    (pv == 1 ? SDNQ_ATTN_PV(HDV_, true, 1) : pv == 2 ? SDNQ_ATTN_PV(HDV_, true, 2) : bf ? SDNQ_ATTN_PV(HDV_, true, 0) : SDNQ_ATTN_PV(HDV_, false, 0))
    return HDV == 128 ? SDNQ_ATTN(128) : SDNQ_ATTN(64);
#undef SDNQ_ATTN_PV
#undef SDNQ_ATTN
}

extern "C" int sdnq_b200_attn_colmean(const void* k, int k_dtype, int64_t heads, int64_t N, int64_t HD, float* mean, void* stream) {
    SDNQ_REQUIRE(k && mean, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(heads >= 0 && heads < (1LL << 31) && N > 0 && N < (1LL << 31), SDNQ_EINVAL, "attn_colmean: bad shape");
    SDNQ_REQUIRE(HD == 16 || HD == 32 || HD == 64 || HD == 128 || HD == 256, SDNQ_EUNSUPPORTED, "attn_colmean: head dim must be 16 .. 256, a power of two (got %lld)", (long long)HD);
    SDNQ_REQUIRE((reinterpret_cast<uintptr_t>(k) & 15) == 0 && (reinterpret_cast<uintptr_t>(mean) & 15) == 0, SDNQ_EINVAL, "attn_colmean: pointers must be 16-byte aligned");
    if (heads == 0) return SDNQ_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const dim3 grid(static_cast<unsigned>(heads), kMeanSplit), block(256);
    if (k_dtype == SDNQ_BF16) SDNQ_CUDA_OK(launch_pdl(attn_colmean_kernel<__nv_bfloat16>, grid, block, 0, st, reinterpret_cast<const __nv_bfloat16*>(k), mean, int(N), int(HD)));
    else if (k_dtype == SDNQ_F16) SDNQ_CUDA_OK(launch_pdl(attn_colmean_kernel<__half>, grid, block, 0, st, reinterpret_cast<const __half*>(k), mean, int(N), int(HD)));
    else if (k_dtype == SDNQ_F32) SDNQ_CUDA_OK(launch_pdl(attn_colmean_kernel<float>, grid, block, 0, st, reinterpret_cast<const float*>(k), mean, int(N), int(HD)));
    else return set_error(SDNQ_EUNSUPPORTED, "attn_colmean: dtype %d", k_dtype);
    return check_launch("attn_colmean_kernel");
}

extern "C" int sdnq_b200_attn_quant(const void* x, int x_dtype, int64_t rows, int64_t HD, const float* mean, int64_t rows_per_head, int mm_dtype,
                                    void* xq, float* scale, void* stream) {
    SDNQ_REQUIRE(x && xq && scale, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(rows >= 0 && (mean == nullptr || rows_per_head > 0), SDNQ_EINVAL, "attn_quant: bad shape");
    SDNQ_REQUIRE(HD == 16 || HD == 32 || HD == 64 || HD == 128 || HD == 256, SDNQ_EUNSUPPORTED, "attn_quant: head dim must be 16 .. 256, a power of two (got %lld)", (long long)HD);
    SDNQ_REQUIRE(mm_dtype == SDNQ_I8 || mm_dtype == SDNQ_F8E4M3, SDNQ_EUNSUPPORTED, "attn_quant: int8 or float8_e4m3fn codes (got %d)", mm_dtype);
    SDNQ_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(xq) | reinterpret_cast<uintptr_t>(mean)) & 15) == 0, SDNQ_EINVAL, "attn_quant: pointers must be 16-byte aligned");
    if (rows == 0) return SDNQ_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t rpw = 32 / (HD / 8), warps = (rows + rpw - 1) / rpw;
    const int64_t want = (warps + 7) / 8, cap = int64_t(num_sms()) * 16;
    const dim3 grid(static_cast<unsigned>(want < cap ? want : cap)), block(256);
    const int rph = int(mean ? rows_per_head : 1);
#define SDNQ_AQ(T, MODE)                                                                                                          \
    do {                                                                                                                          \
        SDNQ_CUDA_OK(launch_pdl(attn_rowquant_kernel<T, MODE>, grid, block, 0, st, reinterpret_cast<const T*>(x), mean, rph, rows, int(HD), \
                                reinterpret_cast<uint8_t*>(xq), scale));                                                          \
        return check_launch("attn_rowquant_kernel");                                                                              \
    } while (0)
    const bool i8 = mm_dtype == SDNQ_I8;
    if (x_dtype == SDNQ_BF16) { if (i8) SDNQ_AQ(__nv_bfloat16, SDNQ_I8); else SDNQ_AQ(__nv_bfloat16, SDNQ_F8E4M3); }
    if (x_dtype == SDNQ_F16) { if (i8) SDNQ_AQ(__half, SDNQ_I8); else SDNQ_AQ(__half, SDNQ_F8E4M3); }
    if (x_dtype == SDNQ_F32) { if (i8) SDNQ_AQ(float, SDNQ_I8); else SDNQ_AQ(float, SDNQ_F8E4M3); }
#undef SDNQ_AQ
    return set_error(SDNQ_EUNSUPPORTED, "attn_quant: dtype %d", x_dtype);
}
