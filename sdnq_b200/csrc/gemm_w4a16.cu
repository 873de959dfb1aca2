// K6: the dequant-path Linear as ONE kernel -- "W4A16": packed weights dequantised in the GEMM prologue, SVD low-rank
// correction as a second accumulate, the bf16 weight never written.
//
// Reference behaviour restated here (use_quantized_matmul=False):
//   quantized_linear_forward            layers/linear/forward.py:24-26    y = F.linear(x, dequantize(W), bias)
//   dequantize_symmetric / _asymmetric  dequantizer.py:15-84              W = cast_T(q * s) | cast_T(fma(q, s, zp)),
//                                                                         then W = cast_T(W + svd_up @ svd_down) for SVD layers
// i.e. per call the reference writes the whole [N,K] weight in the activation dtype and reads it back in a library GEMM
// (2 + 2 + bits/8 bytes of HBM traffic per weight element, two launches).  Here
//   y[m,n] = sum_k x[m,k] * cast_T(q[n,k] * s[n,k/g])  +  sum_j cast_T(sum_k x[m,k] * down[j,k]) * up[n,j]  +  bias[n]
// The first term uses exactly the values the reference's dequantised weight holds before its SVD add; the rank-r term is
// applied to the activations instead of being folded into (and rounded with) every weight: (x W_svd^T) = (x down^T) up^T.
// Mathematically the same Linear; numerically it skips the reference's second rounding of W, so parity is at the stated
// tolerance of every path that contains a 16-bit GEMM (tests/test_layers_gpu.py), not bit-level.
//
// Structure (persistent CTA per SM, 23 warps, warp-specialised; tile = 128 activation rows x 128 weight rows, k-block = 64).
// Three rings of different depth, because their producers have different latencies and their slots different sizes:
//   A ring (x tile 16 KB [+ svd_down tile], 4-6 deep: L2 latency), P ring (packed codes, 4 KB, 8-12 deep: the weights are cold in
//   HBM and a slot is cheap), B ring (dequantised operand tile, 16 KB, 3 deep: produced on chip).  (First version: one 4-deep ring
//   for all three -- 0.6 us per k-block, latency-bound on the weight fetch; profiles/r02_ncu_w4a16_v1.md.)
//   warp 0      TMA producer of the A ring: x tile [128 x 64] (128 B swizzle) and, for SVD layers, the svd_down tile [r x 64] per
//               k-block; the svd_up tile [128 x r] once per tile.  Weight-side loads are issued before griddepcontrol.wait.
//   warp 22     TMA producer of the P ring: packed weight tile [128 rows x 32 B] (linear) per k-block, running far ahead.
//   warps 6-21  dequantise, two groups of 8 warps taking alternate k-blocks: one thread = one weight row x 32 codes (LDS.128 of
//               packed bytes, PRMT nibble -> f32, packed f32x2 subtract / multiply, cvt.rn.bf16x2, 4 x STS.128 into the 128 B-swizzled
//               UMMA operand tile), fence.proxy.async, arrive on the B slot's full barrier; the group scale of the next k-block
//               is fetched one iteration ahead.  (One group of 8 warps was issue-latency-bound: ~1100 cycles per k-block.)
//   warp 1      one thread issues tcgen05.mma kind::f16: acc[128 x 128] += x_tile * Wdq_tile^T (4 per k-block) and, for SVD
//               layers, low[128 x r] += x_tile * down_tile^T into a second TMEM region; after the last k-block the epilogue
//               warps turn `low` into a bf16 operand tile in shared memory and two more MMAs add low * up_tile^T into acc.
//   warps 2-5   epilogue: TMEM -> registers (+ bias) -> bf16 -> swizzled staging -> TMA store (clips M / N tails);
//               two accumulator stages, so the epilogue of tile i overlaps the main loop of tile i+1.
// Bound: tensor pipe at the 16-bit rate for large M; the SD-XL bs=1 shapes are latency-bound (one tile per CTA).
// Algorithmic bytes per launch: N*K/2 (codes) + 4*N*K/g (scales) + 2*r*(N+K) (factors) + 2*M*K (x) + 2*M*N (y).
#include <cstdlib>
#include <mutex>

#include "ptx.cuh"
#include "unpack.cuh"

namespace sdnq {
namespace {

constexpr int BM = 128;            // activation rows per tile = TMEM lanes
constexpr int BN = 128;            // weight rows per tile = accumulator columns
constexpr int BK = 64;             // 16-bit elements of K per stage row: one 128 B swizzle span
constexpr int UMMA_K = 16;
constexpr int kDqWarps = 8;        // warps per dequantise group (one group fills one B tile: 256 threads = 128 rows x 2 halves)
constexpr int kDqGroups = 2;       // the groups take alternate k-blocks: each has two k-block periods for one tile
constexpr int kPWarp = 6 + kDqWarps * kDqGroups;                       // the packed-weight producer warp
constexpr int kThreads = 32 * (kPWarp + 1);                            // 23 warps: <= 88 registers per thread
constexpr int kStoreBlkBytes = 32 * 128;
constexpr int kStoreBytes = 4 * kStoreBlkBytes;                        // one 32-row x 128 B block per epilogue warp: 16 KB
constexpr int kSmemLimit = 227 * 1024;
constexpr int kStageA = BM * 128;                                      // 16 KB
constexpr int kStageB = BN * 128;                                      // 16 KB
constexpr int kStageP = BN * (BK / 2);                                 // 4 KB of 4-bit codes
constexpr int kStagesB = 4;                                            // two slots per dequantise group

template <int RMAX>
struct Cfg {
    static constexpr int kStageD = RMAX * 128;                         // svd_down tile [r x 64] bf16
    static constexpr int kUpBytes = BN * RMAX * 2;                     // svd_up tile [128 x r]
    static constexpr int kLowBytes = BM * RMAX * 2;                    // bf16(x down^T) tile [128 x r]
    static constexpr int kStagesA = RMAX == 0 ? 5 : RMAX <= 32 ? 4 : 3;
    static constexpr int kStagesP = RMAX == 0 ? 8 : 6;                                 // even: a slot always belongs to the same group
    // group scales (and zero points) of the current tile's 128 weight rows, [group][row]: the dequantise loop must not touch
    // global memory (its per-k-block fence would wait for the load: ~700 cycles per k-block in the first versions)
    static constexpr int kScaleBytes = RMAX <= 32 ? 24 * 1024 : 16 * 1024;
    static constexpr int kBarBytes = 8 * (2 * kStagesA + 2 * kStagesB + 2 * kStagesP + 8) + 16;
    static constexpr int kSmemBytes = kStagesA * (kStageA + kStageD) + kStagesB * kStageB + kStagesP * kStageP + kStoreBytes + kUpBytes + kLowBytes +
                                      kScaleBytes + BN * 4 + ((kBarBytes + 127) / 128) * 128;
    static_assert(kSmemBytes <= kSmemLimit, "shared memory budget exceeded");
};

struct W4Args {
    const float* scale;            // [N, groups_per_row]
    const float* zp;               // same shape or NULL
    const void* bias;              // [N] or NULL
    int bias_dtype;
    int M, N, K;
    int group;                     // columns per scale (K for row-wise)
    int group_shift;               // log2(group) when it is a power of two, else -1
    int gpr;                       // groups per row
    int rank;                      // 0 = no SVD term; else 16 / 32 / 64
    float offset;                  // value subtracted from the 4-bit code (8 for int4, 0 for uint4)
    uint32_t fmt16;                // instruction-descriptor operand format: 1 = bf16, 0 = f16
    int dbg;                       // timing experiments only (SDNQ_B200_W4A16_DBG): 1 = no dequantise work, 2 = no proxy fence, 4 = no MMAs
    uint32_t magic;                // 0x4B000000 (2^23 as f32 bits), passed as data so that it lives in a register: PRMT then takes
                                   // the byte selector as its immediate instead of materialising eight selectors per octet
};

template <bool kBf16>
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi) {
    uint32_t r;
    if constexpr (kBf16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// one 32-bit word of packed 4-bit codes (8 values, value i in nibble i) -> 4 registers of activation-dtype pairs holding
// cast_T(q * s) or cast_T(fma(q, s, z))                                                         dequantizer.py:15-84
// `magic` = 0x4B000000 held in a register (so that the byte selectors stay immediates of the PRMTs)
template <bool kBf16, bool kAsym>
__device__ __forceinline__ void dequant8(uint32_t w, uint32_t magic, float2 s2, float2 z2, float2 nb, uint32_t (&out)[4]) {
    const uint32_t lo = w & 0x0F0F0F0Fu, hi = (w >> 4) & 0x0F0F0F0Fu;         // even / odd values, one per byte
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // byte i of lo / hi -> 2^23 + code (exact int -> float without the conversion pipe), then the packed f32x2 pipe
        float2 q = make_float2(__uint_as_float(__byte_perm(lo, magic, 0x7440 | i)), __uint_as_float(__byte_perm(hi, magic, 0x7440 | i)));
        q = __fadd2_rn(q, nb);                                                 // code - offset, exact
        const float2 v = kAsym ? __ffma2_rn(q, s2, z2) : __fmul2_rn(q, s2);
        out[i] = pack16x2<kBf16>(v.x, v.y);
    }
}

template <int RMAX, bool kBf16>
__global__ void __launch_bounds__(kThreads, 1)
gemm_w4a16_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_p,
                  const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_u,
                  const __grid_constant__ CUtensorMap tmap_o, const W4Args a) {
    using C = Cfg<RMAX>;
    constexpr bool kSvd = RMAX > 0;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = ptx::smem_u32(smem_raw);
    if ((smem_base & 1023u) != 0) __trap();
    // [A ring][B ring][D ring (with A)][P ring], then the store blocks, up tile, low tile, bias vector, barriers
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_a + C::kStagesA * kStageA;
    const uint32_t smem_d = smem_b + kStagesB * kStageB;
    const uint32_t smem_p = smem_d + C::kStagesA * C::kStageD;
    const uint32_t smem_o = smem_p + C::kStagesP * kStageP;
    const uint32_t smem_u = smem_o + kStoreBytes;                     // 1024 B aligned: every term above is a multiple of 1 KB
    const uint32_t smem_l = smem_u + C::kUpBytes;
    const uint32_t smem_sc = smem_l + C::kLowBytes;
    float* s_scale = reinterpret_cast<float*>(smem_raw + (smem_sc - smem_base));
    const uint32_t smem_vec = smem_sc + C::kScaleBytes;
    float* s_bias = reinterpret_cast<float*>(smem_raw + (smem_vec - smem_base));
    const uint32_t bar_base = smem_vec + BN * 4;
    auto afull_bar = [&](int s) { return bar_base + 8u * s; };                                       // x (+ down) bytes landed
    auto aempty_bar = [&](int s) { return bar_base + 8u * (C::kStagesA + s); };                      // the k-block's MMAs retired
    auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * C::kStagesA + s); };                   // B tile written (8 warps)
    auto bempty_bar = [&](int s) { return bar_base + 8u * (2 * C::kStagesA + kStagesB + s); };
    auto pfull_bar = [&](int s) { return bar_base + 8u * (2 * C::kStagesA + 2 * kStagesB + s); };    // packed codes landed
    auto pempty_bar = [&](int s) { return bar_base + 8u * (2 * C::kStagesA + 2 * kStagesB + C::kStagesP + s); };
    const uint32_t misc = bar_base + 8u * (2 * C::kStagesA + 2 * kStagesB + 2 * C::kStagesP);
    auto tfull_bar = [&](int s) { return misc + 8u * s; };
    auto tempty_bar = [&](int s) { return misc + 8u * (2 + s); };
    const uint32_t lowfull_bar = misc + 8u * 4;       // low accumulator complete (MMA -> epilogue)
    const uint32_t lowready_bar = misc + 8u * 5;      // bf16 low tile in shared memory (epilogue -> MMA)
    const uint32_t ufull_bar = misc + 8u * 6;         // svd_up tile landed
    const uint32_t ufree_bar = misc + 8u * 7;         // the tile's rank-r MMAs retired: up / low tiles reusable
    const uint32_t tmem_slot = misc + 8u * 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (a.M + BM - 1) / BM, num_n = (a.N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = a.K / BK;
    const int rank = a.rank;
    const uint32_t tx_stage = kStageA + (kSvd ? uint32_t(rank) * 128u : 0u);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_x);
        ptx::prefetch_tmap(&tmap_p);
        ptx::prefetch_tmap(&tmap_o);
        if constexpr (kSvd) {
            ptx::prefetch_tmap(&tmap_d);
            ptx::prefetch_tmap(&tmap_u);
        }
        for (int s = 0; s < C::kStagesA; ++s) {
            ptx::mbar_init(afull_bar(s), 1);
            ptx::mbar_init(aempty_bar(s), 1);
        }
        for (int s = 0; s < kStagesB; ++s) {
            ptx::mbar_init(bfull_bar(s), kDqWarps);
            ptx::mbar_init(bempty_bar(s), 1);
        }
        for (int s = 0; s < C::kStagesP; ++s) {
            ptx::mbar_init(pfull_bar(s), 1);
            ptx::mbar_init(pempty_bar(s), kDqWarps);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), 4);
        }
        ptx::mbar_init(lowfull_bar, 1);
        ptx::mbar_init(lowready_bar, 4);
        ptx::mbar_init(ufull_bar, 1);
        ptx::mbar_init(ufree_bar, 1);
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
    const uint32_t tmem_low = tmem_base + 2 * BN;
    pdl_launch_dependents();

    if (warp == 0) {
        // ======================================================== TMA producer of the A ring (x, svd_down; svd_up per tile)
        if (lane == 0) {
            const int tile0 = blockIdx.x;
            // The weight side never depends on the stream predecessor, the activations do: put the svd tiles of the first ring-full
            // in flight, wait for the predecessor, then add the x tiles of those stages.
            const int npre = tile0 < num_tiles ? (num_kb < C::kStagesA ? num_kb : C::kStagesA) : 0;
            if (npre > 0) {
                if constexpr (kSvd) {
                    ptx::mbar_arrive_expect_tx(ufull_bar, uint32_t(BN) * uint32_t(rank) * 2u);
                    ptx::tma_load_2d(smem_u, &tmap_u, ufull_bar, 0, (tile0 % num_n) * BN);
                }
                for (int s = 0; s < npre; ++s) {
                    ptx::mbar_arrive_expect_tx(afull_bar(s), tx_stage);
                    if constexpr (kSvd) ptx::tma_load_2d(smem_d + s * C::kStageD, &tmap_d, afull_bar(s), s * BK, 0);
                }
            }
            pdl_wait();
            if (npre > 0) {
                const int m0 = (tile0 / num_n) * BM;
                for (int s = 0; s < npre; ++s) ptx::tma_load_2d(smem_a + s * kStageA, &tmap_x, afull_bar(s), s * BK, m0);
            }
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = tile0; tile < num_tiles; tile += gridDim.x, ++it) {
                const int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * BN;
                if constexpr (kSvd) {
                    if (it > 0) {
                        ptx::mbar_wait(ufree_bar, (it - 1) & 1u);          // the previous tile's rank-r MMAs have read the up tile
                        ptx::mbar_arrive_expect_tx(ufull_bar, uint32_t(BN) * uint32_t(rank) * 2u);
                        ptx::tma_load_2d(smem_u, &tmap_u, ufull_bar, 0, n0);
                    }
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (it == 0 && kb < npre) {                            // already in flight (prefetched above)
                        if (++stage == C::kStagesA) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    ptx::mbar_wait(aempty_bar(stage), phase ^ 1u);
                    ptx::mbar_arrive_expect_tx(afull_bar(stage), tx_stage);
                    if constexpr (kSvd) ptx::tma_load_2d(smem_d + stage * C::kStageD, &tmap_d, afull_bar(stage), kb * BK, 0);
                    ptx::tma_load_2d(smem_a + stage * kStageA, &tmap_x, afull_bar(stage), kb * BK, m0);
                    if (++stage == C::kStagesA) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == kPWarp) {
        // ======================================================== TMA producer of the P ring (packed codes), far ahead of the rest
        if (lane == 0) {
            int ps = 0;
            uint32_t pphase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n0 = (tile % num_n) * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(pempty_bar(ps), pphase ^ 1u);
                    ptx::mbar_arrive_expect_tx(pfull_bar(ps), kStageP);
                    ptx::tma_load_2d(smem_p + ps * kStageP, &tmap_p, pfull_bar(ps), kb * (BK / 2), n0);
                    if (++ps == C::kStagesP) { ps = 0; pphase ^= 1u; }
                }
            }
            pdl_wait();
        }
    } else if (warp == 1) {
        // ======================================================== MMA issuer
        if (lane == 0) {
            pdl_wait();
            const uint32_t idesc = ptx::make_idesc(1, a.fmt16, a.fmt16, BM, BN);
            const uint32_t idesc_low = ptx::make_idesc(1, a.fmt16, a.fmt16, BM, kSvd ? uint32_t(rank) : 16u);
            int stage = 0, bs = 0;
            uint32_t phase = 0, bphase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1u;
                ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(afull_bar(stage), phase);
                    ptx::mbar_wait(bfull_bar(bs), bphase);
                    ptx::tc_fence_after();
                    const uint64_t a_desc = ptx::make_smem_desc_sw128(smem_a + stage * kStageA);
                    const uint64_t b_desc = ptx::make_smem_desc_sw128(smem_b + bs * kStageB);
                    if (!(a.dbg & 4)) {
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)      // 16 elements = 32 B along the swizzle span: +2 in >>4 units
                        ptx::umma_f16(d_tmem, a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if constexpr (kSvd) {
                        const uint64_t d_desc = ptx::make_smem_desc_sw128(smem_d + stage * C::kStageD);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            ptx::umma_f16(tmem_low, a_desc + uint64_t(2 * k), d_desc + uint64_t(2 * k), idesc_low, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(aempty_bar(stage));
                    ptx::umma_commit(bempty_bar(bs));
                    if (++stage == C::kStagesA) { stage = 0; phase ^= 1u; }
                    if (++bs == kStagesB) { bs = 0; bphase ^= 1u; }
                }
                if constexpr (kSvd) {
                    ptx::umma_commit(lowfull_bar);                      // x down^T of this tile is complete
                    ptx::mbar_wait(lowready_bar, it & 1u);              // ... and back in shared memory as a bf16 operand tile
                    ptx::mbar_wait(ufull_bar, it & 1u);
                    ptx::tc_fence_after();
                    const uint64_t l_desc = ptx::make_smem_desc_kmajor(smem_l, rank * 2);
                    const uint64_t u_desc = ptx::make_smem_desc_kmajor(smem_u, rank * 2);
                    for (int k = 0; k < rank / UMMA_K; ++k)
                        ptx::umma_f16(d_tmem, l_desc + uint64_t(2 * k), u_desc + uint64_t(2 * k), idesc, 1u);
                    ptx::umma_commit(ufree_bar);
                }
                ptx::umma_commit(tfull_bar(as));
            }
        }
    } else if (warp >= 6 && warp < kPWarp) {
        // ======================================================== dequantise: packed codes -> UMMA operand tile
        // Two groups of 8 warps take alternate k-blocks (group = parity of the CTA's running k-block count), so each has two
        // k-block periods per tile.  thread = (weight row, half of the k-block): 16 packed bytes in, 32 values = 64 B out as
        // four 16 B chunks of the row's 128 B swizzled line (chunk ^ (row & 7)).
        const int grp = (warp - 6) / kDqWarps;
        const int t = threadIdx.x - 192 - grp * (32 * kDqWarps);
        const int row = t & (BN - 1), half = t >> 7;
        const float nbias = -(8388608.0f + a.offset);
        const float2 nb = make_float2(nbias, nbias);
        const uint32_t magic = a.magic;
        const bool asym = a.zp != nullptr;
        auto group_of = [&](int kb) {
            const int k = kb * BK + half * 32;
            return a.gpr > 1 ? (a.group_shift >= 0 ? (k >> a.group_shift) : k / a.group) : 0;
        };
        // Group scales live in shared memory for the duration of a tile ([group][row] floats, zero points behind them) when they
        // fit; otherwise (very long rows / tiny groups) they are fetched from global memory inside the loop, one iteration ahead.
        const bool sc_smem = a.gpr > 1 && num_kb >= 2 && a.gpr * BN * (asym ? 8 : 4) <= C::kScaleBytes;   // (num_kb >= 2: both groups visit every tile)
        float* s_zp = s_scale + a.gpr * BN;
        const int tq = threadIdx.x - 192;                                // 0..511 over both groups
        // ring positions of this group's first k-block (c = grp) and their step of kDqGroups
        int ps = grp, bs = grp;
        uint32_t pphase = 0, bphase = 0;
        int kb = grp, tile = blockIdx.x, tile_sc = -1;
        while (kb >= num_kb && tile < num_tiles) { kb -= num_kb; tile += gridDim.x; }       // (K = 64: a tile is a single k-block)
        while (tile < num_tiles) {
            const int n = (tile % num_n) * BN + row;
            const bool row_ok = n < a.N;
            const float* srow = a.scale + int64_t(row_ok ? n : 0) * a.gpr;
            const float* zrow = asym ? a.zp + int64_t(row_ok ? n : 0) * a.gpr : nullptr;
            if (sc_smem && tile != tile_sc) {
                // both groups are done with the previous tile's scales, then all 512 threads fill the new ones (4 threads per row)
                // (with K = 64 the groups would work on different tiles: such shapes have gpr == 1 and never come here)
                asm volatile("bar.sync 2, 512;" ::: "memory");
                const int r2 = tq & (BN - 1), part = tq >> 7;
                const int n2 = (tile % num_n) * BN + r2;
                for (int g = part; g < a.gpr; g += 4) {
                    s_scale[g * BN + r2] = n2 < a.N ? __ldg(a.scale + int64_t(n2) * a.gpr + g) : 0.f;
                    if (asym) s_zp[g * BN + r2] = n2 < a.N ? __ldg(a.zp + int64_t(n2) * a.gpr + g) : 0.f;
                }
                asm volatile("bar.sync 2, 512;" ::: "memory");
                tile_sc = tile;
            }
            int g_cur = group_of(kb);
            float s, z = 0.f;
            if (sc_smem) {
                s = s_scale[g_cur * BN + row];
                if (asym) z = s_zp[g_cur * BN + row];
            } else {
                s = row_ok ? __ldg(srow + g_cur) : 0.f;
                if (asym) z = row_ok ? __ldg(zrow + g_cur) : 0.f;
            }
            for (; kb < num_kb; kb += kDqGroups) {
                // the scale of this group's next k-block is fetched now and used one iteration later
                int g_nxt = g_cur;
                float s_nxt = s, z_nxt = z;
                if (kb + kDqGroups < num_kb) {
                    g_nxt = group_of(kb + kDqGroups);
                    if (g_nxt != g_cur) {
                        if (sc_smem) {
                            s_nxt = s_scale[g_nxt * BN + row];
                            if (asym) z_nxt = s_zp[g_nxt * BN + row];
                        } else if (row_ok) {
                            s_nxt = __ldg(srow + g_nxt);
                            if (asym) z_nxt = __ldg(zrow + g_nxt);
                        }
                    }
                }
                ptx::mbar_wait(pfull_bar(ps), pphase);                // codes landed
                uint32_t w0, w1, w2, w3;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                             : "r"(smem_p + ps * kStageP + uint32_t(row * 32 + half * 16)));
                const uint32_t words[4] = {w0, w1, w2, w3};
                ptx::mbar_wait(bempty_bar(bs), bphase ^ 1u);          // the B slot's previous MMAs retired
                const uint32_t dst = smem_b + bs * kStageB + uint32_t(row) * 128u;
                const float2 s2 = make_float2(s, s), z2 = make_float2(z, z);
                if (a.dbg & 1) {
                } else if (asym) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t o[4];
                        dequant8<kBf16, true>(words[j], magic, s2, z2, nb, o);
                        ptx::st_shared_v4(dst + (uint32_t((half * 4 + j) ^ (row & 7)) << 4), o[0], o[1], o[2], o[3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t o[4];
                        dequant8<kBf16, false>(words[j], magic, s2, z2, nb, o);
                        ptx::st_shared_v4(dst + (uint32_t((half * 4 + j) ^ (row & 7)) << 4), o[0], o[1], o[2], o[3]);
                    }
                }
                if (!(a.dbg & 2)) ptx::fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(bfull_bar(bs));
                    ptx::mbar_arrive(pempty_bar(ps));
                }
                ps += kDqGroups;
                if (ps >= C::kStagesP) { ps -= C::kStagesP; pphase ^= 1u; }
                bs += kDqGroups;
                if (bs >= kStagesB) { bs -= kStagesB; bphase ^= 1u; }
                s = s_nxt;
                z = z_nxt;
                g_cur = g_nxt;
            }
            do { kb -= num_kb; tile += gridDim.x; } while (kb >= num_kb && tile < num_tiles);
        }
        pdl_wait();
    } else {
        // ======================================================== epilogue (warps 2..5)
        pdl_wait();
        const int q = warp & 3;
        const uint32_t my_o = smem_o + uint32_t(warp - 2) * kStoreBlkBytes;
        int it = 0, blk = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1u;
            const int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * BN;
            const int mrow0 = m0 + q * 32;
            asm volatile("bar.sync 1, 128;" ::: "memory");            // previous tile's readers of s_bias are done
            {
                const int c = threadIdx.x - 64, nc = n0 + c;
                float bv = 0.f;
                if (a.bias != nullptr && nc < a.N)
                    bv = a.bias_dtype == SDNQ_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.bias)[nc])
                         : a.bias_dtype == SDNQ_F16 ? __half2float(reinterpret_cast<const __half*>(a.bias)[nc])
                                                    : reinterpret_cast<const float*>(a.bias)[nc];
                s_bias[c] = bv;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if constexpr (kSvd) {
                // low = x down^T of this tile: TMEM f32 -> activation dtype -> K-major operand tile [128 rows x rank], rows of
                // 2*rank bytes with the swizzle of that span: 16 B chunk c of row m sits at chunk c ^ ((m / (128 / span)) % (span / 16))
                ptx::mbar_wait(lowfull_bar, it & 1u);
                ptx::tc_fence_after();
                const int m = q * 32 + lane;
                const uint32_t span = uint32_t(rank) * 2u;
                const uint32_t xr = (uint32_t(m) / (128u / span)) & (span / 16u - 1u);
                for (int c0 = 0; c0 < rank; c0 += 32) {
                    uint32_t r[32];
                    if (rank - c0 >= 32) {
                        ptx::tmem_ld32(tmem_low + (uint32_t(q * 32) << 16) + c0, r);
                    } else {
                        uint32_t (&r16)[16] = *reinterpret_cast<uint32_t (*)[16]>(&r[0]);
                        ptx::tmem_ld16(tmem_low + (uint32_t(q * 32) << 16) + c0, r16);
                    }
                    ptx::tmem_ld_wait();
                    const int ncols = rank - c0 >= 32 ? 32 : 16;
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        if (8 * c8 < ncols) {
                            uint32_t w[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                w[j] = pack16x2<kBf16>(__uint_as_float(r[8 * c8 + 2 * j]), __uint_as_float(r[8 * c8 + 2 * j + 1]));
                            const uint32_t chunk = uint32_t(c0 / 8 + c8);
                            ptx::st_shared_v4(smem_l + uint32_t(m) * span + ((chunk ^ xr) << 4), w[0], w[1], w[2], w[3]);
                        }
                    }
                }
                ptx::fence_proxy_async_smem();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(lowready_bar);
            }
            ptx::mbar_wait(tfull_bar(as), aphase);
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + as * BN;
#pragma unroll 1
            for (int cb = 0; cb < BN / 64; ++cb) {
                const int n = n0 + cb * 64;
                if (n >= a.N || mrow0 >= a.M) break;                  // warp-uniform
                const uint32_t buf = my_o;
                if (blk > 0) {                                        // the previous bulk store has read the staging block
                    if (lane == 0) ptx::tma_store_wait_read<0>();
                    __syncwarp();
                }
                const uint32_t row_addr = buf + uint32_t(lane) * 128u;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {                         // 32 columns at a time: the kernel lives in 88 registers per thread
                    uint32_t r[32];
                    ptx::tmem_ld32(t_row + cb * 64 + h * 32, r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        const float4 b0 = *reinterpret_cast<const float4*>(s_bias + cb * 64 + h * 32 + 8 * c8);
                        const float4 b1 = *reinterpret_cast<const float4*>(s_bias + cb * 64 + h * 32 + 8 * c8 + 4);
                        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        uint32_t w[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            w[j] = pack16x2<kBf16>(__uint_as_float(r[8 * c8 + 2 * j]) + bv[2 * j], __uint_as_float(r[8 * c8 + 2 * j + 1]) + bv[2 * j + 1]);
                        ptx::st_shared_v4(row_addr + (uint32_t((h * 4 + c8) ^ (lane & 7)) << 4), w[0], w[1], w[2], w[3]);
                    }
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_2d(&tmap_o, buf, n, mrow0);
                    ptx::tma_store_commit();
                }
                ++blk;
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(as));
        }
        if (lane == 0) ptx::tma_store_wait_read<0>();
        __syncwarp();
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

// ------------------------------------------------------------------------------------------------ host
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

// [rows, cols] matrix of `elem_bytes`-wide elements, row pitch `pitch_elems`; box {box_cols, box_rows}; swizzle span in bytes
// (0 = none, else 32 / 64 / 128 = the box row width)
int make_tmap(CUtensorMap* map, const void* ptr, int elem_bytes, int64_t rows, int64_t cols, int64_t pitch_elems, int box_cols, int box_rows, int swizzle) {
    EncodeTiledFn enc = encode_fn();
    SDNQ_REQUIRE(enc != nullptr, SDNQ_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch_elems) * elem_bytes};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(map, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(ptr), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SDNQ_REQUIRE(r == CUDA_SUCCESS, SDNQ_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld pitch=%lld box=%dx%d elem=%d)",
                 static_cast<int>(r), (long long)rows, (long long)cols, (long long)pitch_elems, box_cols, box_rows, elem_bytes);
    return SDNQ_OK;
}

template <int RMAX, bool kBf16>
int launch(const void* x, int64_t ldx, const void* packed, const void* down_rk, const void* up_nr, void* out, const W4Args& a, cudaStream_t st) {
    using C = Cfg<RMAX>;
    auto kernel = gemm_w4a16_kernel<RMAX, kBf16>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes); });
    SDNQ_REQUIRE(attr_err == cudaSuccess, SDNQ_ECUDA, "cudaFuncSetAttribute(max dynamic smem %d) failed: %s", C::kSmemBytes, cudaGetErrorString(attr_err));
    CUtensorMap tx, tp, td, tu, to;
    int rc = make_tmap(&tx, x, 2, a.M, a.K, ldx, BK, BM, 128);
    if (rc != SDNQ_OK) return rc;
    rc = make_tmap(&tp, packed, 1, a.N, a.K / 2, a.K / 2, BK / 2, BN, 0);
    if (rc != SDNQ_OK) return rc;
    rc = make_tmap(&to, out, 2, a.M, a.N, a.N, 64, 32, 128);
    if (rc != SDNQ_OK) return rc;
    if (RMAX > 0) {
        rc = make_tmap(&td, down_rk, 2, a.rank, a.K, a.K, BK, a.rank, 128);
        if (rc != SDNQ_OK) return rc;
        rc = make_tmap(&tu, up_nr, 2, a.N, a.rank, a.rank, a.rank, BN, a.rank * 2);
        if (rc != SDNQ_OK) return rc;
    } else {
        td = tx;
        tu = tx;
    }
    const int tiles = ((a.M + BM - 1) / BM) * ((a.N + BN - 1) / BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    cudaError_t e = launch_pdl(kernel, dim3(grid), dim3(kThreads), C::kSmemBytes, st, tx, tp, td, tu, to, a);
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of gemm_w4a16_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("gemm_w4a16_kernel");
}

}  // namespace
}  // namespace sdnq

using namespace sdnq;

extern "C" int sdnq_b200_linear_w4a16(const void* x, int x_dtype, int64_t ldx, const void* weight, const sdnq_weight_format* fmt,
                                      const float* scale, const float* zero_point, int64_t group_size, const void* svd_down_rk,
                                      const void* svd_up_nr, int svd_rank, const void* bias, int bias_dtype, void* out,
                                      int64_t M, int64_t N, int64_t K, void* stream) {
    SDNQ_REQUIRE(x && weight && scale && out, SDNQ_EINVAL, "NULL pointer");
    WFormat f;
    int rc = make_wformat(fmt, &f);
    if (rc != SDNQ_OK) return rc;
    SDNQ_REQUIRE(f.kind == SDNQ_W_INT && f.bits == 4, SDNQ_EUNSUPPORTED, "linear_w4a16: int4 / uint4 weights (got kind %d, %d bits)", f.kind, f.bits);
    SDNQ_REQUIRE(x_dtype == SDNQ_BF16 || x_dtype == SDNQ_F16, SDNQ_EUNSUPPORTED, "linear_w4a16: bf16 / f16 activations (got %d)", x_dtype);
    SDNQ_REQUIRE(M >= 0 && N > 0 && K > 0 && M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), SDNQ_EINVAL, "bad shape M=%lld N=%lld K=%lld",
                 (long long)M, (long long)N, (long long)K);
    SDNQ_REQUIRE(K % BK == 0 && N % 8 == 0 && ldx % 8 == 0 && ldx >= K, SDNQ_EUNSUPPORTED,
                 "linear_w4a16: K %% 64 == 0, N %% 8 == 0, ldx %% 8 == 0 (K=%lld N=%lld ldx=%lld)", (long long)K, (long long)N, (long long)ldx);
    const int64_t group = (group_size <= 0 || group_size >= K) ? K : group_size;
    SDNQ_REQUIRE(K % group == 0 && (group == K || group % 32 == 0), SDNQ_EUNSUPPORTED,
                 "linear_w4a16: scale groups must be multiples of 32 columns dividing K (group=%lld)", (long long)group);
    SDNQ_REQUIRE(f.is_unsigned == 0 || zero_point != nullptr, SDNQ_EINVAL, "uint4 weights need zero points");
    SDNQ_REQUIRE(svd_rank == 0 || ((svd_rank == 16 || svd_rank == 32 || svd_rank == 64) && svd_down_rk && svd_up_nr), SDNQ_EUNSUPPORTED,
                 "linear_w4a16: svd rank must be 16, 32 or 64 with both factors given (got %d)", svd_rank);
    SDNQ_REQUIRE(bias == nullptr || bias_dtype == SDNQ_BF16 || bias_dtype == SDNQ_F16 || bias_dtype == SDNQ_F32, SDNQ_EINVAL, "bad bias dtype %d", bias_dtype);
    const uintptr_t align = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(out) |
                            reinterpret_cast<uintptr_t>(svd_down_rk) | reinterpret_cast<uintptr_t>(svd_up_nr);
    SDNQ_REQUIRE((align & 15) == 0, SDNQ_EINVAL, "x, weight, out and the svd factors must be 16-byte aligned");
    if (M == 0) return SDNQ_OK;
    int gshift = -1;
    for (int b = 0; b < 31; ++b)
        if ((int64_t(1) << b) == group) gshift = b;
    W4Args a{scale, zero_point, bias, bias_dtype, int(M), int(N), int(K), int(group), gshift, int(K / group), svd_rank,
             f.is_unsigned ? 0.0f : 8.0f, x_dtype == SDNQ_BF16 ? 1u : 0u, getenv("SDNQ_B200_W4A16_DBG") ? atoi(getenv("SDNQ_B200_W4A16_DBG")) : 0,
             0x4B000000u};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool bf = x_dtype == SDNQ_BF16;
    if (svd_rank == 0) return bf ? launch<0, true>(x, ldx, weight, nullptr, nullptr, out, a, st) : launch<0, false>(x, ldx, weight, nullptr, nullptr, out, a, st);
    if (svd_rank <= 32)
        return bf ? launch<32, true>(x, ldx, weight, svd_down_rk, svd_up_nr, out, a, st) : launch<32, false>(x, ldx, weight, svd_down_rk, svd_up_nr, out, a, st);
    return bf ? launch<64, true>(x, ldx, weight, svd_down_rk, svd_up_nr, out, a, st) : launch<64, false>(x, ldx, weight, svd_down_rk, svd_up_nr, out, a, st);
}
