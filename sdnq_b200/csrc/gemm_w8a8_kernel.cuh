// K1: W8A8 scaled matmul on the 5th-gen tensor cores (tcgen05), int8 x int8 -> s32 and e4m3 x e4m3 -> f32.
//
// Reference behaviour restated here:
//   int_scaled_mm_func / fp8_scaled_mm_func    kernel_wrappers.py:193-204
//   sdnq_scaled_mm (Triton)                    kernels/triton_scaled_mm.py:112-275  (acc*sx, then fma(.,sw,bias), cast)
//   int_mm_func / fp8_mm_func -> sdnq_triton_mm kernel_wrappers.py:160-181, kernels/triton_mm.py:13-150
//   zero-point rank-1 terms                    layers/linear/linear_int8.py:65-69, linear_uint8.py:66-73
//
// Structure (one persistent CTA per SM, 6 warps, warp-specialised):
//   warp 0  TMA producer : A tile [128 x 128 B] and B tile [BN x 128 B] per k-block, both K-major, 128 B swizzle,
//                          into a STAGES-deep shared-memory ring (full/empty mbarriers)
//   warp 1  MMA issuer   : one thread issues 4 x tcgen05.mma (K = 32 B each) per k-block into a TMEM accumulator
//                          [128 lanes x BN columns]; two accumulator stages so the epilogue of tile i overlaps
//                          the main loop of tile i+1; tcgen05.commit releases smem slots / publishes the accumulator
//   warps 2-5 epilogue   : tcgen05.ld (lane = output row), f32 epilogue in the reference's operation order,
//                          swizzled shared-memory staging and TMA bulk stores (clip the M / N tails)
// A = activations [M,K] (row-major, K contiguous); B = weight, the reference's K-major [K,N] operand, i.e.
// physically [N,K] with K contiguous -- exactly the "TN" shape tcgen05 wants, so no transposes anywhere.
#pragma once
#include <cmath>
#include <mutex>
#include <type_traits>

#include "act_quant.cuh"
#include "ptx.cuh"
#include "unpack.cuh"

namespace sdnq {
namespace {

constexpr int BM = 128;
constexpr int BK = 128;       // bytes (= elements) of K per stage row: one 128 B swizzle span
constexpr int UMMA_K = 32;    // 8-bit operands: 32 elements per tcgen05.mma
constexpr int kThreads = 192;
constexpr int kEpiWarps = 4;
constexpr int kStoreBufs = 2;                       // per-warp ring of 32-row x 128 B output blocks
constexpr int kStoreBlkBytes = 32 * 128;
constexpr int kStoreBytes = kEpiWarps * kStoreBufs * kStoreBlkBytes;   // 32 KB
constexpr int kSmemLimit = 227 * 1024;

enum OutKind { OUT_BF16 = 0, OUT_F16 = 1, OUT_F32 = 2, OUT_RAW32 = 3 };

}  // namespace

// (named namespace: the one type that crosses between the translation units that instantiate the kernel)
struct GemmParams {
    const float* sx;
    const float* sw;
    const void* bias;
    int bias_dtype;
    int64_t bias_ld;
    const int32_t* rowsum;
    const float* zp;
    const int32_t* colsum;
    const float* zx;
    void* out;
    int out_dtype;
    int M, N, K;
    int raw;   // plain mm: store the accumulator (s32 / f32) untouched
    uint32_t w_sub;   // packed 4-bit weights: per-byte offset removed while expanding (0x08080808 for int4, 0 for uint4)
    uint32_t b_fmt;   // fp8 GEMM: format of the B operand in the instruction descriptor (0 = e4m3, 1 = e5m2); A is always e4m3
    // SVD branch (kSvd): low [M, svd_rank] = cast(x_rot @ svd_down) from svd_low.cu and svd_up as [N, svd_rank], both row-major 16-bit
    const void* svd_low;
    const void* svd_up;
    int svd_rank;     // 16, 32 or 64
    uint32_t svd_fmt; // kind::f16 operand format: 1 = bf16, 0 = f16
    // fused activation quantiser (XM != 0): un-quantised activations in, xq / sx written by this kernel
    const void* fx;
    int64_t fldx;
    uint8_t* fxq;
    float* fsx;
    int* fsync;       // [kSyncStrips] rows quantised per 128-row strip | [kSyncStrips] tiles that consumed the strip
    // packed weights of any width (WB == 7 instantiation): storage bits 2..7, integer codes (pk_kind 0: code - pk_sub per byte) or
    // minifloat codes (pk_kind 1: e<pk_exp>m<pk_man>, expanded to the e4m3 byte of the same value through a 128-entry table)
    int pk_bits, pk_kind, pk_exp, pk_man, pk_unsigned;
    // stream-K (kSK instantiation): parked partial accumulators [CTA][128 x BN] of 32-bit words, and one flag per CTA and epilogue warp
    uint32_t* sk_ws;
    int* sk_flags;
    // grouped launch (n_groups != 0): several Linears fed by the same activations in one launch.  B, sw, bias, zp and colsum are the
    // siblings' tensors concatenated along N, each segment starting at grp_start[g] (a multiple of the tile width) and holding
    // grp_n[g] real rows (zero rows fill the tail); N = grp_start[n_groups]; every sibling has its own contiguous [M, grp_n[g]] output
    int n_groups;
    int grp_start[9];
    int grp_n[8];
    void* grp_out[8];
};

// gemm_w8a8_packed.cu: packed weights of 2 / 3 / 5 / 6 / 7 bits (integer or minifloat codes) expanded in the GEMM prologue
int launch_gemm_packed_any(const void* a, const void* b, const GemmParams& p, cudaStream_t st);

namespace {

constexpr int kSyncStrips = 512;
constexpr int kMaxGroups = 8;
struct OutMaps { CUtensorMap m[kMaxGroups]; };       // output tensor maps: m[0] for a plain launch, one per sibling for a grouped one

// WB = storage bits of the B operand: 8 (int8 / fp8 tiles land in the ring directly by TMA), 7 = any packed width 2..7 given at
// run time (p.pk_bits; the staging ring is sized for 7 bits), or 4 (packed int4 / uint4:
// TMA stages the packed tile, four unpack warps expand it into the ring -- "unpack in the GEMM prologue").
// CG = CTAs per MMA (tcgen05 cta_group): 1, or 2 = a CTA pair computes a 256 x BN tile, each CTA staging its own 128 rows of A
// and only half of the B tile (BN/2 weight rows) -- per SM that is 16 + BN/2 * 128 B per k-block instead of 16 + BN * 128 B.
constexpr int kSvdTileBytes = 128 * 64 * 2;          // one [128 x rank <= 64] 16-bit operand tile

template <int BN, int WB = 8, int CG = 1, bool kSvd = false>
struct Cfg {
    static constexpr int kStageA = BM * BK;
    static constexpr int kStageB = BN / CG * BK;
    static constexpr int kStageBytes = kStageA + kStageB;
    static constexpr int kVecBytes = 2 * BN * 4;                        // sw[BN] and bias[BN] of the current tile as f32
    static constexpr int kPStages = WB < 8 ? 3 : 0;                     // packed staging ring
    static constexpr int kStageP = WB < 8 ? BN * (BK * WB / 8) : 0;
    static constexpr int kThreads = WB < 8 ? 320 : 192;
    static constexpr int kSvdBytes = kSvd ? 2 * kSvdTileBytes : 0;      // low tile [128 x r] + svd_up tile [BN x r], BN = 128
    static constexpr int kLutBytes = WB == 7 ? 128 : 0;                 // minifloat code -> e4m3 byte
    static constexpr int kFixed = kStoreBytes + kVecBytes + kPStages * kStageP + kSvdBytes + 256 /*barriers: 8*(2*stages+4+2*3+2)+4+16 <= 244 B*/ + kLutBytes;
    static constexpr int kStagesRaw = (kSmemLimit - kFixed) / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kAccCols = (kSvd ? 4 : 2) * BN;               // two accumulator stages (+ two f32 stages of the rank-r SVD product)
    static constexpr int kTmemCols = (kAccCols <= 32) ? 32 : (kAccCols <= 64) ? 64 : (kAccCols <= 128) ? 128 : (kAccCols <= 256) ? 256 : 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + kFixed;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128 must be a multiple of 16 in [16,256]");
    static_assert(kStages >= 3, "pipeline too shallow");
    static_assert(CG == 1 || (CG == 2 && WB == 8 && BN % 32 == 0), "CTA pairs: unpacked operands, BN/2 a multiple of 16");
    static_assert(!kSvd || (BN == 128 && CG == 1), "the SVD accumulate runs with 128-wide single-CTA tiles (4 x 128 TMEM columns)");
};

// One row of a packed weight tile (128 weights = 16 octets of BITS bytes, contiguous in shared memory at `src`, 16-byte aligned) ->
// the 128 operand bytes of row `row` of the 128 B-swizzled B stage at `dst`.  kInt: byte = code - sub (offset-binary -> two's
// complement, per byte without cross-byte borrows); else byte = lut[code] (minifloat -> e4m3).
template <int BITS, bool kInt>
__device__ __forceinline__ void expand_packed_row(uint32_t src, uint32_t dst, int row, uint32_t sub, const uint8_t* lut) {
    if constexpr (BITS < 2 || BITS > 7) {
        return;                                                   // (1-bit and 8-bit weights never take this path)
    } else {
        uint32_t w[4 * BITS + 1];
#pragma unroll
        for (int i = 0; i < BITS; ++i)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4 * i]), "=r"(w[4 * i + 1]), "=r"(w[4 * i + 2]), "=r"(w[4 * i + 3]) : "r"(src + 16u * i));
        w[4 * BITS] = 0u;
        constexpr int NW = OctetWords<BITS>::N;
#pragma unroll
        for (int c = 0; c < 8; ++c) {                             // 16-byte chunk c of the row = octets 2c, 2c + 1
            uint32_t o[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                constexpr int kDummy = 0; (void)kDummy;
                const int byte0 = (2 * c + h) * BITS;             // compile-time after unrolling
                uint32_t ow[NW];
#pragma unroll
                for (int i = 0; i < NW; ++i) {
                    const int b = byte0 + 4 * i;
                    ow[i] = (b & 3) ? __funnelshift_r(w[b >> 2], w[(b >> 2) + 1], 8 * (b & 3)) : w[b >> 2];
                }
                uint32_t v[8];
                decode_octet<BITS>(ow, v);
                if constexpr (kInt) {
                    const uint32_t a = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24), b2 = v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24);
                    o[2 * h] = ((a | 0x80808080u) - sub) ^ 0x80808080u;
                    o[2 * h + 1] = ((b2 | 0x80808080u) - sub) ^ 0x80808080u;
                } else {
                    o[2 * h] = uint32_t(lut[v[0]]) | (uint32_t(lut[v[1]]) << 8) | (uint32_t(lut[v[2]]) << 16) | (uint32_t(lut[v[3]]) << 24);
                    o[2 * h + 1] = uint32_t(lut[v[4]]) | (uint32_t(lut[v[5]]) << 8) | (uint32_t(lut[v[6]]) << 16) | (uint32_t(lut[v[7]]) << 24);
                }
            }
            ptx::st_shared_v4(dst + uint32_t(row * 128) + (uint32_t(c ^ (row & 7)) << 4), o[0], o[1], o[2], o[3]);
        }
    }
}

// 8 consecutive values of a f32 / bf16 / f16 vector as floats (16 B aligned for 2-byte types, 32 B for f32)
__device__ __forceinline__ void load8_any(const void* p, int64_t i, int dtype, float (&v)[8]) {
    if (dtype == SDNQ_BF16) load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(p) + i, v);
    else if (dtype == SDNQ_F16) load8<__half>(reinterpret_cast<const __half*>(p) + i, v);
    else load8<float>(reinterpret_cast<const float*>(p) + i, v);
}

// kSimple: epilogue is exactly  out = fma(acc * sx[m], sw[n], bias[n])  (bias[n] = 0 when absent): no zero-point
// terms and no [M,N] bias.  It is the case of every int8 / fp8 symmetric layer and is kept free of the generic
// path's branches so the unrolled epilogue stays small (instruction cache) and at ~4 instructions per element.
//
// XM != 0 (1: bf16 activations, 2: f16): the activation quantiser runs inside this kernel ("phase 1").  Every CTA
// quantises an equal share of the rows of x (1-D bulk copies into the still idle A half of the ring, amax + quantise
// from shared memory, codes and row scales to the workspace in L2), publishes per-128-row-strip progress counters with
// release semantics, and the TMA producer of a tile acquires its strip's counter before it loads A.  One launch per
// Linear instead of two: the dependent-launch gap (~2.5 us, as long as the GEMM itself at SD-XL sizes) disappears,
// and the weight prefetch overlaps the quantisation.  All CTAs are co-resident (grid <= SMs, 1 CTA/SM), which the
// cross-CTA wait relies on.
// kSvd: the SVD branch of the W8A8 forwards (linear_int8.py:57-62).  The reference adds bias2d = bias + (x @ svd_down) @ svd_up as
// a dense [M,N] bias; here the rank-r product low[128 x r] . svd_up[BN x r]^T of the tile is one more tcgen05.mma (kind::f16) after
// the k-loop into its own f32 TMEM region, and the epilogue adds it to the bias in f32 before the fma.
// kSK (stream-K): the launch's k-blocks (tile-major: tile 0's k-blocks, tile 1's, ...) are split EVENLY over the CTAs instead of
// tile by tile, so a GEMM of 80 tiles keeps all 148 SMs loading and multiplying.  A CTA's share is a run of segments (tile, k-block
// range).  A segment that starts inside a tile parks its raw accumulator in global memory and raises a flag; the CTA whose segment
// holds the tile's FIRST k-block -- by construction the last segment of its run, while every parked piece of that tile is the first
// segment of a higher-numbered CTA's run -- adds the parked pieces to its own accumulator and runs the epilogue.  Nobody waits for a
// CTA that could be waiting itself; all CTAs are co-resident (grid = SMs, one CTA per SM).  int8 only: 32-bit integer partial sums
// add up to exactly the accumulator of the un-split tile, so the output is bit-identical.
template <int BN, bool kInt8, int OUT, bool kSimple, int WB, int XM, int CG, bool kSvd = false, bool kSK = false>
__global__ void __launch_bounds__((Cfg<BN, WB, CG, kSvd>::kThreads), 1)
gemm_w8a8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ OutMaps tmaps_o, const __grid_constant__ CUtensorMap tmap_l,
                 const __grid_constant__ CUtensorMap tmap_u, const GemmParams p) {
    using C = Cfg<BN, WB, CG, kSvd>;
    static_assert(!kSvd || (XM == 0 && !kSimple), "SVD tiles use the generic epilogue and the stand-alone activation quantiser");
    static_assert(CG == 1 || XM == 0, "the fused quantiser runs with single-CTA MMAs");
    static_assert(!kSK || (kInt8 && CG == 1 && WB == 8 && XM == 0 && !kSvd && OUT != OUT_RAW32), "stream-K: plain int8 single-CTA tiles");
    constexpr bool kPair = CG == 2;
    // CTA pair: rank 0 (the leader) issues the MMAs; tiles are numbered per pair (256 rows x BN columns)
    const uint32_t cta_rank = kPair ? ptx::cluster_ctarank() : 0u;
    const int tile_first = int(blockIdx.x) / CG, tile_step = int(gridDim.x) / CG;
    constexpr bool kPacked = WB < 8;
    const int pk_bits = WB == 7 ? p.pk_bits : WB;                      // storage bits per weight; one tile row of 128 weights = 16 * pk_bits bytes
    constexpr int kOutBytes = (OUT == OUT_BF16 || OUT == OUT_F16) ? 2 : 4;
    constexpr int CPB = 128 / kOutBytes;              // output columns per 128 B store block: 64 or 32
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = ptx::smem_u32(smem_raw);
    if ((smem_base & 1023u) != 0) __trap();          // the 128 B swizzle atoms need a 1024 B aligned base
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_base + C::kStages * C::kStageA;
    const uint32_t smem_o = smem_base + C::kStages * C::kStageBytes;
    float* s_sw = reinterpret_cast<float*>(smem_raw + C::kStages * C::kStageBytes + kStoreBytes);
    float* s_bias = s_sw + BN;
    const uint32_t smem_p = smem_o + kStoreBytes + C::kVecBytes;       // packed B staging ring (kPacked only)
    const uint32_t smem_l = smem_p + C::kPStages * C::kStageP;         // kSvd: low tile, then the svd_up tile (1 KB aligned: kVecBytes = 1 KB at BN = 128)
    const uint32_t smem_u = smem_l + kSvdTileBytes;
    const uint32_t bar_base = smem_l + C::kSvdBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + 2 + s); };
    auto pfull_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + 4 + s); };
    auto pempty_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + 4 + C::kPStages + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 4 + 2 * C::kPStages);
    const uint32_t xstage_bar = tmem_slot + 8u;                        // fused quantiser: bulk copies of x landed
    const uint32_t svdfull_bar = tmem_slot + 16u;                      // kSvd: low / svd_up tiles landed
    const uint32_t svdfree_bar = tmem_slot + 24u;                      // kSvd: the tile's rank-r MMAs retired
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (p.M + BM * CG - 1) / (BM * CG), num_n = (p.N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = (p.K + BK - 1) / BK;
    auto tile_m0 = [&](int tile) { return (tile / num_n) * (BM * CG) + int(cta_rank) * BM; };      // this CTA's 128 output rows
    auto tile_nb0 = [&](int tile) { return (tile % num_n) * BN + int(cta_rank) * (BN / CG); };      // this CTA's share of the B rows
    // the barrier TMA completions are counted on: the leader's (it alone waits for the operands of both CTAs)
    auto load_bar = [&](int s) { return kPair ? ptx::mapa(full_bar(s), 0) : full_bar(s); };
    auto tma_load = [&](uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
        if constexpr (kPair) ptx::tma_load_2d_pair(dst, tmap, bar, c0, c1);
        else ptx::tma_load_2d(dst, tmap, bar, c0, c1);
    };
    auto expect_stage = [&](int s) { if (cta_rank == 0) ptx::mbar_arrive_expect_tx(full_bar(s), C::kStageBytes * CG); };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        ptx::prefetch_tmap(&tmaps_o.m[0]);
        for (int s = 0; s < C::kStages; ++s) {
            ptx::mbar_init(full_bar(s), kPacked ? 1 + 4 : 1);     // TMA (A) [+ the four unpack warps (B)]
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), kEpiWarps * CG);       // pair: the epilogue warps of both CTAs release the leader's MMA warp
        }
        for (int s = 0; s < C::kPStages; ++s) {
            ptx::mbar_init(pfull_bar(s), 1);
            ptx::mbar_init(pempty_bar(s), 4);
        }
        if constexpr (XM != 0) ptx::mbar_init(xstage_bar, 1);
        if constexpr (kSvd) {
            ptx::prefetch_tmap(&tmap_l);
            ptx::prefetch_tmap(&tmap_u);
            ptx::mbar_init(svdfull_bar, 1);
            ptx::mbar_init(svdfree_bar, 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (kPair) { ptx::tmem_alloc_pair(tmem_slot, C::kTmemCols); ptx::tmem_relinquish_pair(); }
        else { ptx::tmem_alloc(tmem_slot, C::kTmemCols); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before();
    if constexpr (kPair) ptx::cluster_sync();     // the peer's barriers are initialised before anything signals them
    else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_launch_dependents();     // our own dependents may begin their prologue

    // ---- work of this CTA as a run of segments (tile, k-blocks [kb0, kb1)): whole tiles, or (kSK) an even share of all k-blocks
    struct Seg { int tile, kb0, kb1; };
    const int64_t sk_total = int64_t(num_tiles) * num_kb;
    auto sk_first_unit = [&](int cta) { return static_cast<int>(sk_total * cta / int(gridDim.x)); };
    const int sk_u0 = kSK ? sk_first_unit(blockIdx.x) : 0, sk_u1 = kSK ? sk_first_unit(blockIdx.x + 1) : 0;
    const int seg_start = kSK ? sk_u0 : tile_first;
    auto seg_next = [&](int& cursor, Seg& sg) -> bool {
        if constexpr (kSK) {
            if (cursor >= sk_u1) return false;
            sg.tile = cursor / num_kb;
            sg.kb0 = cursor - sg.tile * num_kb;
            sg.kb1 = min(num_kb, sg.kb0 + (sk_u1 - cursor));
            cursor += sg.kb1 - sg.kb0;
        } else {
            if (cursor >= num_tiles) return false;
            sg.tile = cursor;
            sg.kb0 = 0;
            sg.kb1 = num_kb;
            cursor += tile_step;
        }
        return true;
    };
    // weight prefetch bookkeeping shared by the producer branch below: the first segment's first k-blocks
    const int tile0 = kSK ? sk_u0 / num_kb : tile_first;
    const int kb_first = kSK ? sk_u0 - tile0 * num_kb : 0;
    const int first_len = kSK ? min(num_kb - kb_first, sk_u1 - sk_u0) : num_kb;
    const bool has_work = kSK ? sk_u0 < sk_u1 : tile0 < num_tiles;
    const int npre = (!kPacked && has_work) ? (first_len < C::kStages ? first_len : C::kStages) : 0;
    if constexpr (XM != 0) {
        // ======================================================== phase 1: quantise this CTA's share of the activation rows
        if (warp == 0 && lane == 0 && npre > 0) {                 // weights first: they do not depend on anything
            const int n0 = (tile0 % num_n) * BN;
            for (int s = 0; s < npre; ++s) {
                ptx::mbar_arrive_expect_tx(full_bar(s), C::kStageBytes);
                ptx::tma_load_2d(smem_b + s * C::kStageB, &tmap_b, full_bar(s), s * BK, n0);
            }
        }
        pdl_wait();                                               // x belongs to the stream predecessor
        using XT = typename std::conditional<XM == 1, __nv_bfloat16, __half>::type;
        constexpr int MODE = kInt8 ? SDNQ_I8 : SDNQ_F8E4M3;
        constexpr int kWarpsAll = C::kThreads / 32;
        const int K = p.K;
        const int rpc = (p.M + int(gridDim.x) - 1) / int(gridDim.x);
        const int r0 = min(p.M, int(blockIdx.x) * rpc), r1 = min(p.M, r0 + rpc);
        const uint32_t row_bytes = uint32_t(K) * 2u;
        const int rpb = int(uint32_t(C::kStages * C::kStageA) / row_bytes);      // rows per staging batch (host guarantees >= 1)
        uint32_t xphase = 0;
        for (int rb = r0; rb < r1; rb += rpb) {
            const int re = min(r1, rb + rpb);
            if (warp == 0) {
                if (lane == 0) ptx::mbar_arrive_expect_tx(xstage_bar, uint32_t(re - rb) * row_bytes);
                __syncwarp();
                for (int r = rb + lane; r < re; r += 32)
                    ptx::bulk_load_1d(smem_a + uint32_t(r - rb) * row_bytes,
                                      reinterpret_cast<const uint8_t*>(p.fx) + int64_t(r) * p.fldx * 2, row_bytes, xstage_bar);
            }
            ptx::mbar_wait(xstage_bar, xphase);
            xphase ^= 1u;
            for (int r = rb + warp; r < re; r += kWarpsAll) {
                const uint8_t* srow = smem_raw + size_t(r - rb) * row_bytes;
                float amax = 0.f;
#pragma unroll 4
                for (int k = lane * 8; k < K; k += 256) {
                    actq::Held<XT> h;
                    h.raw = *reinterpret_cast<const uint4*>(srow + 2 * k);
                    float v[8];
                    h.get(v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) amax = fmaxf(amax, fabsf(v[i]));
                }
                amax = warp_max(amax);
                const float scale = __fdiv_rn(amax, kInt8 ? 127.f : 448.f);     // get_scale_symmetric (quant_utils.py:264-299)
                const actq::RowDivider divider(scale);
                const bool safe = divider.safe();
                uint8_t* dst = p.fxq + int64_t(r) * K;
                int unused = 0;
#pragma unroll 2
                for (int k = lane * 8; k < K; k += 256) {
                    actq::Held<XT> h;
                    h.raw = *reinterpret_cast<const uint4*>(srow + 2 * k);
                    float v[8];
                    h.get(v);
                    const uint2 q = safe ? actq::quantise8<MODE, true>(v, divider, 0.f, false, unused)
                                         : actq::quantise8<MODE, false>(v, divider, 0.f, false, unused);
                    *reinterpret_cast<uint2*>(dst + k) = q;
                }
                if (lane == 0) p.fsx[r] = scale;
                // publish the row: every lane orders its stores against the async proxy (the TMA engines of the other
                // CTAs read them), the warp converges, one lane bumps the strip counter with release semantics
                ptx::fence_proxy_async_all();
                __syncwarp();
                if (lane == 0) ptx::red_release_gpu_add(p.fsync + r / BM, 1);
            }
            if (rb + rpb < r1) __syncthreads();                   // the staging rows are overwritten by the next batch
        }
        __syncthreads();                                          // staging reads are done before A tiles land in the same memory
    }
    // producer side of the cross-CTA dependency: all rows of the tile's 128-row strip are quantised
    auto acquire_strip = [&](int ms) {
        if constexpr (XM != 0) {
            const int target = min(BM, p.M - ms * BM);
            int spins = 0;
            while (ptx::ld_acquire_gpu(p.fsync + ms) < target) {
                if (++spins > (1 << 27)) __trap();                // seconds: a CTA of this grid never ran (not co-resident?)
            }
            ptx::fence_proxy_async_all();
        }
    };
    // ... and, off the critical path, once the tile's loads are on their way: the last of the strip's num_n consumers
    // re-arms both counters for the next launch on this workspace
    auto release_strip = [&](int ms) {
        if constexpr (XM != 0) {
            if (atomicAdd(p.fsync + kSyncStrips + ms, 1) == num_n - 1) {
                p.fsync[kSyncStrips + ms] = 0;
                p.fsync[ms] = 0;
            }
        }
    };

    // kSvd: the producer fetches the tile's low / svd_up operand tiles once the previous tile's rank-r MMAs have read the old ones
    auto load_svd_tiles = [&](int it, int m0, int n0) {
        if constexpr (kSvd) {
            if (it > 0) ptx::mbar_wait(svdfree_bar, (it - 1) & 1u);
            ptx::mbar_arrive_expect_tx(svdfull_bar, uint32_t(BM + BN) * uint32_t(p.svd_rank) * 2u);
            ptx::tma_load_2d(smem_l, &tmap_l, svdfull_bar, 0, m0);
            ptx::tma_load_2d(smem_u, &tmap_u, svdfull_bar, 0, n0);
        }
    };

    if (warp == 0) {
        // ======================================================== TMA producer
        if (lane == 0 && kPacked) {
            // packed weights: A tiles go to the ring, packed B tiles to the staging ring (the unpack warps fill the ring's B half)
            pdl_wait();
            int stage = 0, ps = 0, pit = 0;
            uint32_t phase = 0, pphase = 0;
            for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++pit) {
                const int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * BN;
                load_svd_tiles(pit, m0, n0);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    ptx::mbar_arrive_expect_tx(full_bar(stage), C::kStageA);
                    ptx::tma_load_2d(smem_a + stage * C::kStageA, &tmap_a, full_bar(stage), kb * BK, m0);
                    ptx::mbar_wait(pempty_bar(ps), pphase ^ 1u);
                    ptx::mbar_arrive_expect_tx(pfull_bar(ps), uint32_t(BN * 16 * pk_bits));
                    ptx::tma_load_2d(smem_p + ps * C::kStageP, &tmap_b, pfull_bar(ps), kb * 16 * pk_bits, n0);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                    if (++ps == C::kPStages) { ps = 0; pphase ^= 1u; }
                }
            }
        } else if (lane == 0) {
            // The weight operand never depends on the stream predecessor (weights are frozen), the activations do: start
            // the first ring-full of B tiles now, then wait for the predecessor (the activation quantiser) and add the A tiles.
            if constexpr (XM == 0) {
                if (npre > 0) {
                    const int nb0 = tile_nb0(tile0);
                    for (int s = 0; s < npre; ++s) {
                        expect_stage(s);
                        tma_load(smem_b + s * C::kStageB, &tmap_b, load_bar(s), (kb_first + s) * BK, nb0);
                    }
                }
                pdl_wait();
            }
            if (npre > 0) {
                const int m0 = tile_m0(tile0);
                acquire_strip(tile0 / num_n);
                for (int s = 0; s < npre; ++s) tma_load(smem_a + s * C::kStageA, &tmap_a, load_bar(s), (kb_first + s) * BK, m0);
            }
            int stage = 0, pit = 0, cursor = seg_start;
            uint32_t phase = 0;
            Seg sg;
            for (; seg_next(cursor, sg); ++pit) {
                const int tile = sg.tile;
                const int m0 = tile_m0(tile), nb0 = tile_nb0(tile);
                if (pit != 0) acquire_strip(tile / num_n);
                load_svd_tiles(pit, m0, (tile % num_n) * BN);
                for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
                    if (pit == 0 && kb < sg.kb0 + npre) {         // already in flight (prefetched above)
                        if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    expect_stage(stage);
                    tma_load(smem_a + stage * C::kStageA, &tmap_a, load_bar(stage), kb * BK, m0);
                    tma_load(smem_b + stage * C::kStageB, &tmap_b, load_bar(stage), kb * BK, nb0);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                }
                release_strip(tile / num_n);
            }
        }
    } else if (warp == 1) {
        // ======================================================== MMA issuer
        if (lane == 0 && cta_rank == 0) {
            pdl_wait();
            // fp8: A = e4m3 activations, B = e4m3 or e5m2 weights (the mixed pair torch._scaled_mm takes for float8_e5m2 weights)
            const uint32_t idesc = kInt8 ? ptx::make_idesc(2, 1, 1, BM * CG, BN) : ptx::make_idesc(1, 0, p.b_fmt, BM * CG, BN);
            auto commit = [&](uint32_t bar) { if constexpr (kPair) ptx::umma_commit_pair(bar); else ptx::umma_commit(bar); };
            int stage = 0;
            uint32_t phase = 0;
            int it = 0, cursor = seg_start;
            Seg sg;
            for (; seg_next(cursor, sg); ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1u;
                ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);     // epilogue has drained this accumulator stage
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);       // TMA bytes have landed
                    ptx::tc_fence_after();
                    const uint64_t a_desc = ptx::make_smem_desc_sw128(smem_a + stage * C::kStageA);
                    const uint64_t b_desc = ptx::make_smem_desc_sw128(smem_b + stage * C::kStageB);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advancing K inside the 128 B swizzle span = advancing the start address (>>4 units)
                        if constexpr (kPair)
                            ptx::umma_ss_pair<kInt8>(d_tmem, a_desc + uint64_t(k * UMMA_K >> 4), b_desc + uint64_t(k * UMMA_K >> 4), idesc,
                                                     (kb != sg.kb0 || k != 0) ? 1u : 0u);
                        else
                            ptx::umma_ss<kInt8>(d_tmem, a_desc + uint64_t(k * UMMA_K >> 4), b_desc + uint64_t(k * UMMA_K >> 4), idesc,
                                                (kb != sg.kb0 || k != 0) ? 1u : 0u);
                    }
                    commit(empty_bar(stage));                     // smem slot free (in both CTAs of a pair) once these MMAs retire
                    if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                }
                if constexpr (kSvd) {
                    ptx::mbar_wait(svdfull_bar, it & 1u);
                    ptx::tc_fence_after();
                    const uint32_t idesc_svd = ptx::make_idesc(1, p.svd_fmt, p.svd_fmt, BM, BN);
                    const uint64_t l_desc = ptx::make_smem_desc_kmajor(smem_l, p.svd_rank * 2);
                    const uint64_t u_desc = ptx::make_smem_desc_kmajor(smem_u, p.svd_rank * 2);
                    for (int k = 0; k < p.svd_rank / 16; ++k)
                        ptx::umma_f16(tmem_base + 2 * BN + as * BN, l_desc + uint64_t(2 * k), u_desc + uint64_t(2 * k), idesc_svd, k != 0 ? 1u : 0u);
                    commit(svdfree_bar);
                }
                commit(tfull_bar(as));                            // accumulator complete
            }
        }
    } else if (kPacked && warp >= 6) {
        // ======================================================== unpack warps (6..9): packed int4 / uint4 -> int8 tile
        // One task = one 16 B chunk of the 128 B-swizzled B tile (16 values of one weight row) = 8 packed bytes: LDS.64,
        // nibble split + byte interleave (PRMT), offset-binary -> two's complement without cross-byte borrows, STS.128.
        // Both sides are bank-conflict free (a warp reads 256 contiguous bytes and writes 4 swizzled 128 B rows).
        const int t = threadIdx.x - 192;
        [[maybe_unused]] const uint8_t* s_lut = smem_raw + (bar_base - smem_base) + 256;
        if constexpr (WB == 7 && !kInt8) {
            // minifloat code -> the e4m3 byte of the same value (every eXmY with X <= 4, Y <= 3 is a subset of e4m3; host checks)
            const float v = decode_minifloat(uint32_t(t), p.pk_bits, p.pk_exp, p.pk_man, p.pk_unsigned);
            smem_raw[(bar_base - smem_base) + 256 + t] = static_cast<uint8_t>(__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3));
            asm volatile("bar.sync 2, 128;" ::: "memory");
        }
        int stage = 0, ps = 0;
        uint32_t phase = 0, pphase = 0;
        for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
            for (int kb = 0; kb < num_kb; ++kb) {
                ptx::mbar_wait(pfull_bar(ps), pphase);            // packed tile landed
                ptx::mbar_wait(empty_bar(stage), phase ^ 1u);     // the ring slot's previous MMAs retired
                const uint32_t src = smem_p + ps * C::kStageP, dst = smem_b + stage * C::kStageB;
                if constexpr (WB == 7) {
                    // any width: thread t expands weight row t of the tile -- its 16 * bits packed bytes (aligned 16-byte loads) hold 16
                    // octets of 8 codes (unpack.cuh), each octet becomes 8 operand bytes of the 128 B-swizzled row
                    static_assert(BN == 128, "one unpack thread per tile row");
                    SDNQ_DISPATCH_BITS(pk_bits, (expand_packed_row<BITS, kInt8>(src + uint32_t(t) * uint32_t(16 * BITS), dst, t, p.w_sub, s_lut)));
                } else
#pragma unroll 4
                for (int i = 0; i < BN * 8 / 128; ++i) {
                    const int idx = i * 128 + t;
                    const int row = idx >> 3, chunk = idx & 7;
                    uint32_t x0, x1;
                    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(x0), "=r"(x1) : "r"(src + uint32_t(row * 64 + chunk * 8)));
                    uint32_t o[4];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t x = h ? x1 : x0;
                        const uint32_t lo = x & 0x0F0F0F0Fu, hi = (x >> 4) & 0x0F0F0F0Fu;     // even / odd nibbles
                        const uint32_t a = __byte_perm(lo, hi, 0x5140), b = __byte_perm(lo, hi, 0x7362);   // values 0-3, 4-7
                        o[2 * h] = ((a | 0x80808080u) - p.w_sub) ^ 0x80808080u;                // code - 8 per byte (int4) / code (uint4)
                        o[2 * h + 1] = ((b | 0x80808080u) - p.w_sub) ^ 0x80808080u;
                    }
                    ptx::st_shared_v4(dst + uint32_t(row * 128) + (uint32_t(chunk ^ (row & 7)) << 4), o[0], o[1], o[2], o[3]);
                }
                ptx::fence_proxy_async_smem();                    // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(full_bar(stage));
                    ptx::mbar_arrive(pempty_bar(ps));
                }
                if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                if (++ps == C::kPStages) { ps = 0; pphase ^= 1u; }
            }
        }
    } else {
        // ======================================================== epilogue (warps 2..5)
        // TMEM lane = output row.  Each warp turns its 32 rows x CPB columns into one 32 x 128 B block in shared
        // memory (128 B swizzle, conflict-free 16 B stores) and hands it to a TMA store, which clips the M / N tails
        // and writes full lines; a 2-deep ring per warp overlaps the store with the next block's math.
        pdl_wait();                                               // sx / rowsum / bias may come from the predecessor
        const int q = warp & 3;                                   // TMEM lane quarter this warp may access
        const uint32_t my_o = smem_o + uint32_t(warp - 2) * (kStoreBufs * kStoreBlkBytes);
        int it = 0, blk = 0, cursor = seg_start;
        Seg sg;
        for (; seg_next(cursor, sg); ++it) {
            const int tile = sg.tile;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1u;
            const int m0 = tile_m0(tile), n0 = (tile % num_n) * BN;
            // stream-K: 0 = the segment is the whole tile, 1 = it holds the tile's first k-block (collects the parked pieces and
            // finishes the tile), 2 = it starts inside the tile (parks its accumulator)
            [[maybe_unused]] const int sk_kind = !kSK ? 0 : sg.kb0 != 0 ? 2 : sg.kb1 != num_kb ? 1 : 0;
            [[maybe_unused]] int sk_last = int(blockIdx.x);       // kind 1: CTAs blockIdx.x + 1 .. sk_last parked a piece of this tile
            if constexpr (kSK) {
                if (sk_kind == 1) {
                    const int tile_end = (tile + 1) * num_kb;
                    while (sk_last + 1 < int(gridDim.x) && sk_first_unit(sk_last + 1) < tile_end) ++sk_last;
                }
            }
            // grouped launch: the sibling this tile belongs to; its real columns end at n_end, its output starts at column n_base
            int grp = 0;
            while (grp + 1 < p.n_groups && n0 >= p.grp_start[grp + 1]) ++grp;
            const int n_base = p.n_groups != 0 ? p.grp_start[grp] : 0;
            const int n_end = p.n_groups != 0 ? n_base + p.grp_n[grp] : p.N;
            const CUtensorMap* tmap_o = &tmaps_o.m[grp];
            const int mrow0 = m0 + q * 32;
            const int m = mrow0 + lane;
            const bool m_ok = m < p.M;
            float sxm = 0.f, zxm = 0.f, rsx = 0.f;
            if (XM == 0 && OUT != OUT_RAW32 && m_ok) {
                sxm = p.sx[m];
                if (p.zx) zxm = p.zx[m];
                if (p.rowsum) rsx = __fmul_rn(static_cast<float>(p.rowsum[m]), sxm);   // (rowsum -> f32) * sx
            }
            const bool vec_bias = p.bias != nullptr && p.bias_ld == 0;
            if constexpr (OUT != OUT_RAW32) {
                // per-column vectors of this tile -> shared memory once (f32), read back as broadcast LDS.128;
                // issued before the accumulator wait so the global latency hides behind the main loop
                asm volatile("bar.sync 1, 128;" ::: "memory");            // previous tile's readers are done
                for (int c = threadIdx.x - 64; c < BN; c += 128) {
                    const int nc = n0 + c;
                    float sv = 0.f, bv = 0.f;
                    if (nc < n_end) {
                        sv = p.sw[nc];
                        if (vec_bias)
                            bv = p.bias_dtype == SDNQ_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[nc])
                                 : p.bias_dtype == SDNQ_F16 ? __half2float(reinterpret_cast<const __half*>(p.bias)[nc])
                                                            : reinterpret_cast<const float*>(p.bias)[nc];
                    }
                    s_sw[c] = sv;
                    s_bias[c] = bv;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            ptx::mbar_wait(tfull_bar(as), aphase);
            ptx::tc_fence_after();
            if (XM != 0 && m_ok) sxm = __ldcg(p.sx + m);          // written by phase 1 of some CTA of this grid: read it from L2, after the accumulator
            const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + as * BN;
            if constexpr (kSK) {
                if (sk_kind == 1) {                               // the parked pieces of this warp's rows have landed (whatever the rows hold)
                    if (lane == 0) {
                        for (int c2 = int(blockIdx.x) + 1; c2 <= sk_last; ++c2) {
                            int spins = 0;
                            while (ptx::ld_acquire_gpu(p.sk_flags + c2 * 4 + q) == 0) {
                                if (++spins > (1 << 27)) __trap();    // seconds: a CTA of this grid never ran (not co-resident?)
                            }
                        }
                    }
                    __syncwarp();
                }
            }
#pragma unroll 1
            for (int cb = 0; cb < BN / CPB; ++cb) {
                const int n = n0 + cb * CPB;
                if (n >= n_end || mrow0 >= p.M) break;            // warp-uniform
                uint32_t r[CPB];
                {
                    uint32_t (&r0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[0]);
                    ptx::tmem_ld32(t_row + cb * CPB, r0);
                    if constexpr (CPB == 64) {
                        uint32_t (&r1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[32]);
                        ptx::tmem_ld32(t_row + cb * CPB + 32, r1);
                    }
                }
                const uint32_t buf = my_o + uint32_t(blk % kStoreBufs) * kStoreBlkBytes;
                if (blk >= kStoreBufs) {                          // ring slot still being read by an older TMA store?
                    if (lane == 0) ptx::tma_store_wait_read<kStoreBufs - 1>();
                    __syncwarp();
                }
                ptx::tmem_ld_wait();
                if constexpr (kSK) {
                    // parked pieces: [CTA][warp quarter][column block][16-byte chunk][lane] -- every access is a coalesced 512-byte row
                    auto piece = [&](int cta) {
                        return reinterpret_cast<uint4*>(p.sk_ws + size_t(cta) * (BM * BN)) + (size_t(q) * (BN / CPB) + cb) * (CPB / 4) * 32 + lane;
                    };
                    if (sk_kind == 2) {
                        uint4* dst = piece(int(blockIdx.x));
#pragma unroll
                        for (int i = 0; i < CPB / 4; ++i) dst[i * 32] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
                        continue;
                    }
                    if (sk_kind == 1) {
                        for (int c2 = int(blockIdx.x) + 1; c2 <= sk_last; ++c2) {
                            const uint4* src = piece(c2);
#pragma unroll
                            for (int i = 0; i < CPB / 4; ++i) {
                                const uint4 v = __ldcg(src + i * 32);
                                r[4 * i] += v.x; r[4 * i + 1] += v.y; r[4 * i + 2] += v.z; r[4 * i + 3] += v.w;
                            }
                        }
                    }
                }
                const uint32_t row_addr = buf + uint32_t(lane) * 128u;
                if constexpr (OUT == OUT_RAW32) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        ptx::st_shared_v4(row_addr + (uint32_t(c ^ (lane & 7)) << 4), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
                } else {
                    constexpr int kChunkCols = 16 / kOutBytes;    // 8 (2-byte out) or 4 (f32 out)
#pragma unroll
                    for (int c8 = 0; c8 < CPB / 8; ++c8) {        // groups of 8 columns
                        const int nc = n + 8 * c8;
                        float y[8];
                        if (nc < n_end) {
                            float swv[8];
                            {
                                const float4 s0 = *reinterpret_cast<const float4*>(s_sw + cb * CPB + 8 * c8);
                                const float4 s1 = *reinterpret_cast<const float4*>(s_sw + cb * CPB + 8 * c8 + 4);
                                swv[0] = s0.x; swv[1] = s0.y; swv[2] = s0.z; swv[3] = s0.w;
                                swv[4] = s1.x; swv[5] = s1.y; swv[6] = s1.z; swv[7] = s1.w;
                            }
                            float b[8];
                            if constexpr (kSimple) {
                                const float4 b0 = *reinterpret_cast<const float4*>(s_bias + cb * CPB + 8 * c8);
                                const float4 b1 = *reinterpret_cast<const float4*>(s_bias + cb * CPB + 8 * c8 + 4);
                                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
                                b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const uint32_t rv = r[8 * c8 + j];
                                    const float acc = kInt8 ? static_cast<float>(static_cast<int>(rv)) : __uint_as_float(rv);
                                    y[j] = fmaf(__fmul_rn(acc, sxm), swv[j], b[j]);      // fma(acc*sx, sw, bias)
                                }
                            } else {
                                bool has_b = false;
                                if (p.zp) {                           // zero_bias = (rowsum*sx)*zp            linear_int8.py:66
                                    float zpv[8];
                                    load8<float>(p.zp + nc, zpv);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) b[j] = __fmul_rn(rsx, zpv[j]);
                                    has_b = true;
                                    if (p.colsum) {                   // += (colsum*sw)*zx ; += K*(zx*zp)       linear_uint8.py:67-72
#pragma unroll
                                        for (int j = 0; j < 8; ++j) {
                                            const float wt = __fmul_rn(__fmul_rn(static_cast<float>(p.colsum[nc + j]), swv[j]), zxm);
                                            b[j] = __fadd_rn(b[j], wt);
                                            b[j] = fmaf(static_cast<float>(p.K), __fmul_rn(zxm, zpv[j]), b[j]);
                                        }
                                    }
                                } else if (p.colsum) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) b[j] = __fmul_rn(__fmul_rn(static_cast<float>(p.colsum[nc + j]), swv[j]), zxm);
                                    has_b = true;
                                }
                                if constexpr (kSvd) {                 // + (low @ svd_up^T)[m, n]: the rank-r MMA's f32 accumulator
                                    uint32_t sv[8];
                                    ptx::tmem_ld8(t_row + 2 * BN + cb * CPB + 8 * c8, sv);
                                    ptx::tmem_ld_wait();
                                    if (has_b) {
#pragma unroll
                                        for (int j = 0; j < 8; ++j) b[j] = __fadd_rn(b[j], __uint_as_float(sv[j]));
                                    } else {
#pragma unroll
                                        for (int j = 0; j < 8; ++j) b[j] = __uint_as_float(sv[j]);
                                        has_b = true;
                                    }
                                }
                                if (p.bias && (p.bias_ld == 0 || m_ok)) {
                                    float bv[8];
                                    if (vec_bias) {
                                        const float4 b0 = *reinterpret_cast<const float4*>(s_bias + cb * CPB + 8 * c8);
                                        const float4 b1 = *reinterpret_cast<const float4*>(s_bias + cb * CPB + 8 * c8 + 4);
                                        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
                                        bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
                                    } else {
                                        load8_any(p.bias, int64_t(m) * p.bias_ld + nc, p.bias_dtype, bv);
                                    }
                                    if (has_b) {
#pragma unroll
                                        for (int j = 0; j < 8; ++j) b[j] = __fadd_rn(b[j], bv[j]);
                                    } else {
#pragma unroll
                                        for (int j = 0; j < 8; ++j) b[j] = bv[j];
                                        has_b = true;
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const uint32_t rv = r[8 * c8 + j];
                                    const float acc = kInt8 ? static_cast<float>(static_cast<int>(rv)) : __uint_as_float(rv);
                                    const float t = __fmul_rn(acc, sxm);          // acc * sx
                                    y[j] = has_b ? fmaf(t, swv[j], b[j]) : __fmul_rn(t, swv[j]);
                                }
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) y[j] = 0.f;
                        }
                        if constexpr (kChunkCols == 8) {
                            uint32_t w[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if constexpr (OUT == OUT_BF16) {
                                    __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
                                    w[j] = *reinterpret_cast<uint32_t*>(&h);
                                } else {
                                    __half2 h = __floats2half2_rn(y[2 * j], y[2 * j + 1]);
                                    w[j] = *reinterpret_cast<uint32_t*>(&h);
                                }
                            }
                            ptx::st_shared_v4(row_addr + (uint32_t(c8 ^ (lane & 7)) << 4), w[0], w[1], w[2], w[3]);
                        } else {
                            ptx::st_shared_v4(row_addr + (uint32_t((2 * c8) ^ (lane & 7)) << 4), __float_as_uint(y[0]), __float_as_uint(y[1]),
                                              __float_as_uint(y[2]), __float_as_uint(y[3]));
                            ptx::st_shared_v4(row_addr + (uint32_t((2 * c8 + 1) ^ (lane & 7)) << 4), __float_as_uint(y[4]), __float_as_uint(y[5]),
                                              __float_as_uint(y[6]), __float_as_uint(y[7]));
                        }
                    }
                }
                ptx::fence_proxy_async_smem();                    // generic-proxy smem writes -> visible to the TMA engine
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_2d(tmap_o, buf, n - n_base, mrow0);
                    ptx::tma_store_commit();
                }
                ++blk;
            }
            if constexpr (kSK) {
                if (sk_kind == 2) {                               // publish this warp's parked rows
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) ptx::red_release_gpu_add(p.sk_flags + int(blockIdx.x) * 4 + q, 1);
                } else if (sk_kind == 1) {                        // re-arm the flags for the next launch on this workspace
                    __syncwarp();
                    if (lane == 0)
                        for (int c2 = int(blockIdx.x) + 1; c2 <= sk_last; ++c2) p.sk_flags[c2 * 4 + q] = 0;
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (kPair) ptx::mbar_arrive_cluster(ptx::mapa(tempty_bar(as), 0));
                else ptx::mbar_arrive(tempty_bar(as));
            }
        }
        if (lane == 0) ptx::tma_store_wait_read<0>();             // smem must outlive the last bulk stores
        __syncwarp();
    }
    // ---- teardown
    ptx::tc_fence_before();
    if constexpr (kPair) ptx::cluster_sync();     // neither CTA's shared memory / barriers may go away while the peer can still touch them
    else __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tc_fence_after();
        if constexpr (kPair) ptx::tmem_dealloc_pair(tmem_base, C::kTmemCols);
        else ptx::tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------ host
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// [rows, cols] row-major matrix of `elem_bytes`-wide elements -> 2-D tensor map with a box of {128 B, box_rows},
// 128 B swizzle; loads zero-fill out-of-bounds elements, stores clip them.
int make_tmap(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int elem_bytes, int box_rows, int box_bytes = 128) {
    EncodeTiledFn enc = get_encode_fn();
    SDNQ_REQUIRE(enc != nullptr, SDNQ_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                   : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * elem_bytes};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_bytes / elem_bytes), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    // 128 B boxes are swizzled (UMMA operands, conflict-free staging); narrower boxes (packed weights) stay linear
    CUresult r = enc(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     box_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SDNQ_REQUIRE(r == CUDA_SUCCESS, SDNQ_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld elem=%d box_rows=%d)",
                 static_cast<int>(r), (long long)rows, (long long)cols, elem_bytes, box_rows);
    return SDNQ_OK;
}

// [rows, rank] 16-bit matrix (row pitch = rank) -> tensor map with a box of {rank, box_rows} and the swizzle of a rank*2-byte row
int make_tmap_svd(CUtensorMap* map, const void* ptr, int64_t rows, int rank, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    SDNQ_REQUIRE(enc != nullptr, SDNQ_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(rank), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(rank) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(rank), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = rank == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : rank == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SDNQ_REQUIRE(r == CUDA_SUCCESS, SDNQ_ECUDA, "cuTensorMapEncodeTiled (svd operand) failed with CUresult %d (rows=%lld rank=%d)", static_cast<int>(r), (long long)rows, rank);
    return SDNQ_OK;
}

template <int BN, bool kInt8, int OUT, bool kSimple, int WB = 8, int XM = 0, int CG = 1, bool kSvd = false, bool kSK = false>
int launch_gemm(const void* a, const void* b, const GemmParams& p, cudaStream_t st) {
    using C = Cfg<BN, WB, CG, kSvd>;
    auto kernel = gemm_w8a8_kernel<BN, kInt8, OUT, kSimple, WB, XM, CG, kSvd, kSK>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes); });
    SDNQ_REQUIRE(attr_err == cudaSuccess, SDNQ_ECUDA, "cudaFuncSetAttribute(max dynamic smem %d) failed: %s", C::kSmemBytes,
                 cudaGetErrorString(attr_err));
    CUtensorMap ta, tb;
    OutMaps to;
    int rc = make_tmap(&ta, a, p.M, p.K, 1, BM);
    if (rc != SDNQ_OK) return rc;
    const int pbits = WB == 7 ? p.pk_bits : WB;
    rc = WB < 8 ? make_tmap(&tb, b, p.N, int64_t(p.K) * pbits / 8, 1, BN, BK * pbits / 8) : make_tmap(&tb, b, p.N, p.K, 1, BN / CG);
    if (rc != SDNQ_OK) return rc;
    constexpr int kOutElem = (OUT == OUT_BF16 || OUT == OUT_F16) ? 2 : 4;
    if (p.n_groups == 0) {
        rc = make_tmap(&to.m[0], p.out, p.M, p.N, kOutElem, 32);
        if (rc != SDNQ_OK) return rc;
    } else {
        for (int g = 0; g < p.n_groups; ++g) {
            SDNQ_REQUIRE(p.grp_start[g] % BN == 0, SDNQ_EINVAL, "grouped launch: segment %d starts at %d, not a multiple of the tile width %d", g, p.grp_start[g], BN);
            rc = make_tmap(&to.m[g], p.grp_out[g], p.M, p.grp_n[g], kOutElem, 32);
            if (rc != SDNQ_OK) return rc;
        }
    }
    CUtensorMap tl = ta, tu = ta;
    if (kSvd) {
        rc = make_tmap_svd(&tl, p.svd_low, p.M, p.svd_rank, BM);
        if (rc != SDNQ_OK) return rc;
        rc = make_tmap_svd(&tu, p.svd_up, p.N, p.svd_rank, BN);
        if (rc != SDNQ_OK) return rc;
    }
    const int tiles = ((p.M + BM * CG - 1) / (BM * CG)) * ((p.N + BN - 1) / BN);      // CG == 2: 256-row tiles, one per CTA pair
    // fused quantiser: always one CTA per SM (CTAs without a tile still quantise their share of the rows)
    const int slots = num_sms() / CG;
    const int grid = CG * ((XM != 0 || kSK || tiles >= slots) ? slots : tiles);      // stream-K: every SM gets an even share of the k-blocks
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(C::kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (CG == 2) {                        // the two CTAs of a pair are co-scheduled on one TPC
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = 2;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        cfg.numAttrs = 2;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, ta, tb, to, tl, tu, p);
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of gemm_w8a8_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("gemm_w8a8_kernel");
}

}  // namespace
}  // namespace sdnq
