// K3s: dequant-only path WITH the SVD low-rank correction, rank-r term on the tensor cores.
//
//   W[n,k] = cast( cast(q[n,k] * scale[n,g] (+ zp)) + sum_j svd_up[n,j] * svd_down[j,k] )          reference dequantizer.py:52-84
//
// The rank-r update is a [N x r] . [r x K] GEMM with r = 16 / 32 / 64: 32 FMA per output element on CUDA cores make the
// kernel compute-bound (~2.4x the HBM time), one tcgen05.mma (kind::f16, bf16 operands, f32 accumulate in TMEM) per
// 128 x 256 output tile makes it free.  Structure = the W8A8 GEMM's (persistent CTA per SM, TMA producer warp, one MMA
// thread, 4 epilogue warps, double-buffered TMEM accumulator, swizzled smem staging + TMA bulk stores) with a single
// k-block per tile and a different epilogue: TMEM lane = weight row n, so each epilogue lane dequantises 64 consecutive k of
// its own row per block (packed bytes for the whole 256-wide tile row are fetched up front: 128 B per lane for 4-bit codes,
// 16 KB in flight per SM), adds the accumulator and rounds exactly where the reference's `result.to(svd dtype).addmm_()` does.
//
// Operands: A = svd_up [N, r] (r contiguous: K-major), B = svd_down as stored for the dequant path, logical [r, K] with stride
// (1, r) = physical [K, r] (r contiguous: K-major).  One TMA box row is r*2 bytes = the swizzle span (32 / 64 / 128 B).
#include <cstdlib>
#include <mutex>
#include <new>

#include "ptx.cuh"
#include "unpack.cuh"

namespace sdnq {

namespace {

constexpr int TM = 128;         // weight rows per tile (TMEM lanes)
// TN = weight columns per tile (MMA N): 256, 128 or 64.  The epilogue (dequantise + add + round + store, 4 warps) is the long pole
// of a tile; the SD-XL weights are small (0.4 - 13 M elements), so the tile is narrowed until every SM has one: 128 x 256 tiles
// left a 1280 x 1280 weight on 50 SMs (7.7 us, the side stream could not hide it behind the 6 us GEMM of the previous layer).
constexpr int kThreads = 192;
constexpr int kStages = 2;
// output staging blocks per epilogue warp: a 64-wide tile is one block per warp, so a second buffer would only cost the shared memory
// that lets a 4th CTA live on the SM
template <int TN> constexpr int store_bufs() { return TN == 64 ? 1 : 2; }
constexpr int kStoreBlkBytes = 32 * 128;
constexpr int kMaxRank = 64;
// operand stages are sized by the layer's rank at launch: A = svd_up tile [128 x r], B = svd_down tile [TN x r] (r * 2 bytes per row)
__host__ __device__ constexpr int stage_a_bytes(int rank) { return TM * rank * 2; }
template <int TN> struct SvdCfg {
    __host__ __device__ static constexpr int stage_b_bytes(int rank) { return TN * rank * 2; }
    static constexpr int kUnits = TN / 32;                            // 16 B units of 4-bit codes per tile row
    static constexpr int kPkBytes = kUnits * 32 * 16;                 // one tile row of codes per lane: units x 32 lanes x 16 B
    static constexpr int kPkTotal = 4 * 2 * kPkBytes;                 // 4 epilogue warps x double buffer
    static constexpr int kBlocks = TN / 64;                           // 64-column store blocks per tile
    static constexpr int kScFloats = 2 * kBlocks * 32;                // (scale, zp) x blocks x 32 lanes, per warp and buffer
    static constexpr int kScTotal = 4 * 2 * kScFloats * 4;
    static constexpr int kStoreBufs = store_bufs<TN>();
    static constexpr int kStoreBytes = 4 * kStoreBufs * kStoreBlkBytes;
    static constexpr int kFixedBytes = kStoreBytes + kPkTotal + kScTotal + 256;
    __host__ __device__ static constexpr int smem_bytes(int rank) { return kStages * (stage_a_bytes(rank) + stage_b_bytes(rank)) + kFixedBytes; }
    // Two accumulator stages of TN columns: allocating exactly that (not all 512 columns) and keeping the CTA small lets several
    // CTAs share an SM, which is what hides the per-tile latency chain (TMA -> MMA -> TMEM -> epilogue)
    static constexpr int kTmemCols = 2 * TN;
    static constexpr int kCtasPerSm = 512 / kTmemCols < 4 ? 512 / kTmemCols : 4;      // (launch bound: registers for up to 4 CTAs)
    static int ctas_per_sm(int rank) {
        const int by_smem = (227 * 1024) / (smem_bytes(rank) + 1024);
        return by_smem < 1 ? 1 : by_smem < kCtasPerSm ? by_smem : kCtasPerSm;
    }
};

struct SvdArgs {
    const uint8_t* weight;
    const float* scale;
    const float* zp;
    int N, K;
    int group32, group_shift, gpr32, row_stride32;
    WFormat f;
    int rank;
    int out_dtype;
};

// One weight of a launch: its three tensor maps, its arguments and where its tiles start in the launch's tile numbering.  A single
// dequantisation passes one entry as a kernel parameter; a batched launch (several layers' weights dequantised ahead of their
// GEMMs by one persistent grid) reads a table of them from global memory.
struct alignas(128) BatchEntry {
    CUtensorMap up, down, out;
    SvdArgs a;
    int tile_start;      // index of this weight's first tile in the launch
    int num_n;           // tiles along K
};

// ZP / BLK: what the launch's weights have in common, decided on the host so that the element loop carries no per-weight branches --
// ZP = 0 no weight has zero points, 1 every weight has, 2 mixed (decided per tile);  BLK = 1 every weight's scale groups cover whole
// 64-column blocks (row-wise scales or groups that are multiples of 64: the scales are staged per block), 0 decided per tile.
template <int BITS, bool kBf16Out, int TN, int ZP, int BLK>
__global__ void __launch_bounds__(kThreads, SvdCfg<TN>::kCtasPerSm)
dequant_svd_kernel(const __grid_constant__ BatchEntry single, const BatchEntry* __restrict__ table, const int n_entries, const int total_tiles,
                   const int stage_rank) {
    using SC = SvdCfg<TN>;
    constexpr int kPkBytes = SC::kPkBytes, kPkTotal = SC::kPkTotal, kScTotal = SC::kScTotal, kStoreBufs = SC::kStoreBufs, kStoreBytes = SC::kStoreBytes;
    const BatchEntry* const tab = table != nullptr ? table : &single;
    const int kStageA = stage_a_bytes(stage_rank), kStageB = SC::stage_b_bytes(stage_rank);      // stages sized for the largest rank of the launch
    // tile t of the launch -> the entry it belongs to (t only grows in every role's loop, so the search resumes at `li`)
    auto locate = [&](int t, int& li) -> const BatchEntry* {
        while (li + 1 < n_entries && t >= tab[li + 1].tile_start) ++li;
        return tab + li;
    };
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = ptx::smem_u32(smem_raw);
    if ((smem_base & 1023u) != 0) __trap();
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_base + kStages * kStageA;
    const uint32_t smem_o = smem_base + kStages * (kStageA + kStageB);
    const uint32_t smem_pk = smem_o + kStoreBytes;
    float* s_scales = reinterpret_cast<float*>(smem_raw + kStages * (kStageA + kStageB) + kStoreBytes + kPkTotal);
    const uint32_t bar_base = smem_pk + kPkTotal + kScTotal;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = total_tiles;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, SC::kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int li = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const BatchEntry* e = locate(tile, li);
                const int local = tile - e->tile_start, num_n = e->num_n;
                const int m0 = (local / num_n) * TM, n0 = (local % num_n) * TN;
                ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                ptx::mbar_arrive_expect_tx(full_bar(stage), static_cast<uint32_t>((TM + TN) * e->a.rank * 2));
                ptx::tma_load_2d(smem_a + stage * kStageA, &e->up, full_bar(stage), 0, m0);
                ptx::tma_load_2d(smem_b + stage * kStageB, &e->down, full_bar(stage), 0, n0);
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc(1 /*f32 acc*/, 1 /*bf16*/, 1 /*bf16*/, TM, TN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0, li = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int rank = locate(tile, li)->a.rank;
                const int row_bytes = rank * 2;                        // = swizzle span
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1u;
                ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
                ptx::mbar_wait(full_bar(stage), phase);
                ptx::tc_fence_after();
                const uint64_t a_desc = ptx::make_smem_desc_kmajor(smem_a + stage * kStageA, row_bytes);
                const uint64_t b_desc = ptx::make_smem_desc_kmajor(smem_b + stage * kStageB, row_bytes);
                for (int k = 0; k < rank / 16; ++k)                   // 16 bf16 = 32 B per MMA along the contraction
                    ptx::umma_f16(tmem_base + as * TN, a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc, k != 0 ? 1u : 0u);
                ptx::umma_commit(empty_bar(stage));
                ptx::umma_commit(tfull_bar(as));
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ======================================================== epilogue: dequant + add + round + store
        // Packed codes of a lane's tile row (256 four-bit codes = 128 B) and its per-64-column scales are prefetched one tile
        // ahead into shared memory with cp.async (16 KB in flight per SM, lane-interleaved 16 B units: conflict-free both ways),
        // and consumed by a *rolled* loop over 32-column slices -- the unrolled form of this epilogue overflowed the
        // instruction cache (ncu: 56 % icc hit rate, no_inst stalls), the rolled one is ~400 SASS instructions.
        const int q = warp & 3, ew = warp - 2;
        const uint32_t my_o = smem_o + uint32_t(ew) * (kStoreBufs * kStoreBlkBytes);
        const uint32_t pk_base = smem_pk + uint32_t(ew) * (2 * kPkBytes);              // [2 buffers][8 units][32 lanes][16 B]
        float* sc_base = s_scales + ew * (2 * SC::kScFloats);                           // [2 buffers][scale|zp][blocks][32 lanes]
        // per-64-column scales can be staged ahead only when a scale group never ends inside a 64-column block
        auto blk_scales_of = [](const SvdArgs& a) { return a.gpr32 <= 1 || (a.group32 & 63) == 0; };

        auto prefetch = [&](int tile, int buf, int& li) {
            const BatchEntry* e = locate(tile, li);
            const SvdArgs& a = e->a;
            const bool blk_scales = BLK == 1 || blk_scales_of(a);
            const int local = tile - e->tile_start, num_n = e->num_n;
            const int m0 = (local / num_n) * TM, n0 = (local % num_n) * TN;
            const int n = m0 + q * 32 + lane;
            const bool row_ok = n < a.N;
#pragma unroll
            for (int u = 0; u < SC::kUnits; ++u) {
                const int k = n0 + 32 * u;
                const bool ok = row_ok && k < a.K;
                const uint8_t* src = a.weight + (ok ? ((static_cast<uint32_t>(n) * static_cast<uint32_t>(a.K) + k) >> 1) : 0u);
                const uint32_t dst = pk_base + uint32_t(buf) * kPkBytes + uint32_t(u * 32 + lane) * 16u;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (blk_scales) {
#pragma unroll
                for (int cb = 0; cb < SC::kBlocks; ++cb) {
                    const int k = n0 + cb * 64;
                    float sc = 0.f, z = 0.f;
                    if (row_ok && k < a.K) {
                        const int g = a.gpr32 <= 1 ? 0 : (a.group_shift >= 0 ? (k >> a.group_shift) : static_cast<int>(static_cast<uint32_t>(k) / static_cast<uint32_t>(a.group32)));
                        const uint32_t si = static_cast<uint32_t>(n) * static_cast<uint32_t>(a.row_stride32) + g;
                        sc = a.scale[si];
                        if (ZP == 1 || (ZP == 2 && a.zp)) z = a.zp[si];
                    }
                    sc_base[(buf * 2 + 0) * (SC::kBlocks * 32) + cb * 32 + lane] = sc;
                    sc_base[(buf * 2 + 1) * (SC::kBlocks * 32) + cb * 32 + lane] = z;
                }
            }
        };

        int it = 0, blk = 0, li = 0, li_next = 0;
        if (blockIdx.x < num_tiles) prefetch(blockIdx.x, 0, li_next);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1, buf = it & 1;
            const uint32_t aphase = (it >> 1) & 1u;
            const BatchEntry* e = locate(tile, li);
            const SvdArgs a = e->a;                                        // this tile's weight (registers)
            const CUtensorMap* tmap_out = &e->out;
            const float bias = 8388608.0f - static_cast<float>(a.f.int_offset);
            const bool blk_scales = BLK == 1 || blk_scales_of(a);
            const bool has_zp = ZP == 1 || (ZP == 2 && a.zp != nullptr);
            const int local = tile - e->tile_start, num_n = e->num_n;
            const int m0 = (local / num_n) * TM, n0 = (local % num_n) * TN;
            const int mrow0 = m0 + q * 32;
            const int n = mrow0 + lane;
            const bool row_ok = n < a.N;
            const int next = tile + gridDim.x;
            if (next < num_tiles) {
                prefetch(next, buf ^ 1, li_next);
                asm volatile("cp.async.wait_group 1;" ::: "memory");      // this tile's codes have landed (own lane's data only)
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            ptx::mbar_wait(tfull_bar(as), aphase);
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + as * TN;
            const uint32_t pk = pk_base + uint32_t(buf) * kPkBytes + uint32_t(lane) * 16u;
            const float* scs = sc_base + (buf * 2) * (SC::kBlocks * 32) + lane;
#pragma unroll 1
            for (int cb = 0; cb < TN / 64; ++cb) {
                const int kb0 = n0 + cb * 64;
                if (kb0 >= a.K || mrow0 >= a.N) break;                     // warp-uniform
                const uint32_t sbuf = my_o + uint32_t(blk % kStoreBufs) * kStoreBlkBytes;
                if (blk >= kStoreBufs) {
                    if (lane == 0) ptx::tma_store_wait_read<kStoreBufs - 1>();
                    __syncwarp();
                }
                const uint32_t row_addr = sbuf + uint32_t(lane) * 128u;
                float sc = scs[cb * 32], z = scs[SC::kBlocks * 32 + cb * 32];
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {                              // two 32-column slices per store block
                    const int u = cb * 2 + h;
                    uint32_t r[32];
                    ptx::tmem_ld32(t_row + u * 32, r);
                    uint32_t w0, w1, w2, w3;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(pk + uint32_t(u) * 512u));
                    const uint32_t words[4] = {w0, w1, w2, w3};
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int o = 0; o < 4; ++o) {                          // octet o of the slice
                        const int k = kb0 + h * 32 + o * 8;
                        if (!blk_scales && row_ok && k < a.K) {
                            const int g = a.group_shift >= 0 ? (k >> a.group_shift) : static_cast<int>(static_cast<uint32_t>(k) / static_cast<uint32_t>(a.group32));
                            const uint32_t si = static_cast<uint32_t>(n) * static_cast<uint32_t>(a.row_stride32) + g;
                            sc = a.scale[si];
                            if (has_zp) z = a.zp[si];
                        }
                        const uint32_t lo = words[o] & 0x0F0F0F0Fu, hi = (words[o] >> 4) & 0x0F0F0F0Fu;
                        uint32_t w4[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {                              // byte j of the word = elements 2j (low nibble), 2j + 1
                            // the element pair goes through the packed f32x2 forms (FADD2 / FMUL2 / FFMA2: IEEE per component, one issue
                            // slot for both -- this epilogue is issue-bound)
                            const float2 q = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7440 | j)),
                                                                    __uint_as_float(__byte_perm(hi, 0x4B000000u, 0x7440 | j))), make_float2(-bias, -bias));
                            const float2 v = has_zp ? __ffma2_rn(q, make_float2(sc, sc), make_float2(z, z)) : __fmul2_rn(q, make_float2(sc, sc));
                            // result.to(svd dtype): one packed conversion rounds both to bf16, two bit operations bring them back to f32
                            __nv_bfloat162 wb = __floats2bfloat162_rn(v.x, v.y);
                            const uint32_t wbits = *reinterpret_cast<uint32_t*>(&wb);
                            const float2 y = __fadd2_rn(make_float2(__uint_as_float(wbits << 16), __uint_as_float(wbits & 0xFFFF0000u)),      // addmm_: f32 accumulate,
                                                        make_float2(__uint_as_float(r[8 * o + 2 * j]), __uint_as_float(r[8 * o + 2 * j + 1])));   // rounded once below
                            __nv_bfloat162 hh = __floats2bfloat162_rn(y.x, y.y);
                            w4[j] = *reinterpret_cast<uint32_t*>(&hh);
                        }
                        const int c8 = h * 4 + o;
                        ptx::st_shared_v4(row_addr + (uint32_t(c8 ^ (lane & 7)) << 4), w4[0], w4[1], w4[2], w4[3]);
                    }
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_2d(tmap_out, sbuf, kb0, mrow0);
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
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, SC::kTmemCols);
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

// 2-byte element matrix [rows, cols] with row pitch `pitch_elems`, box {box_cols, box_rows}, swizzle = box_cols * 2 bytes
int make_tmap16(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t pitch_elems, int box_cols, int box_rows, int swizzle_bytes) {
    EncodeTiledFn enc = encode_fn();
    SDNQ_REQUIRE(enc != nullptr, SDNQ_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch_elems) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SDNQ_REQUIRE(r == CUDA_SUCCESS, SDNQ_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld pitch=%lld box=%dx%d)", static_cast<int>(r),
                 (long long)rows, (long long)cols, (long long)pitch_elems, box_cols, box_rows);
    return SDNQ_OK;
}

// fills one table entry (tensor maps, arguments, tile range) for tile width TN
template <int TN>
int fill_entry(BatchEntry* e, const SvdArgs& a, const void* up, int64_t up_pitch, const void* down, int64_t down_pitch, void* out, int tile_start) {
    int rc = make_tmap16(&e->up, up, a.N, a.rank, up_pitch, a.rank, TM, a.rank * 2);
    if (rc != SDNQ_OK) return rc;
    rc = make_tmap16(&e->down, down, a.K, a.rank, down_pitch, a.rank, TN, a.rank * 2);
    if (rc != SDNQ_OK) return rc;
    rc = make_tmap16(&e->out, out, a.N, a.K, a.K, 64, 32, 128);
    if (rc != SDNQ_OK) return rc;
    e->a = a;
    e->tile_start = tile_start;
    e->num_n = (a.K + TN - 1) / TN;
    return SDNQ_OK;
}

template <int TN, int ZP, int BLK>
int launch_variant(const BatchEntry* single, const BatchEntry* device_table, int n_entries, int total_tiles, int stage_rank, cudaStream_t st) {
    constexpr int kSmemMax = SvdCfg<TN>::smem_bytes(kMaxRank);
    const int kSmemBytes = SvdCfg<TN>::smem_bytes(stage_rank);
    auto kernel = dequant_svd_kernel<4, true, TN, ZP, BLK>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax); });
    SDNQ_REQUIRE(attr_err == cudaSuccess, SDNQ_ECUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
    const int slots = num_sms() * SvdCfg<TN>::ctas_per_sm(stage_rank);
    const int grid = total_tiles < slots ? total_tiles : slots;
    static const BatchEntry kNone{};
    cudaError_t e = launch_pdl(kernel, dim3(grid), dim3(kThreads), kSmemBytes, st, single != nullptr ? *single : kNone, device_table, n_entries,
                               total_tiles, stage_rank);
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of dequant_svd_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("dequant_svd_kernel");
}

// traits: bit 0 = some weight has zero points, bit 1 = some weight has none, bit 2 = some weight's scale groups end inside a 64-column block
template <int TN>
int launch_entries(const BatchEntry* single, const BatchEntry* device_table, int n_entries, int total_tiles, int stage_rank, int traits, cudaStream_t st) {
    if ((traits & 4) == 0 && (traits & 3) == 2) return launch_variant<TN, 0, 1>(single, device_table, n_entries, total_tiles, stage_rank, st);
    if ((traits & 4) == 0 && (traits & 3) == 1) return launch_variant<TN, 1, 1>(single, device_table, n_entries, total_tiles, stage_rank, st);
    return launch_variant<TN, 2, 0>(single, device_table, n_entries, total_tiles, stage_rank, st);
}

int entry_traits(const SvdArgs& a) { return (a.zp != nullptr ? 1 : 2) | ((a.gpr32 <= 1 || (a.group32 & 63) == 0) ? 0 : 4); }

bool svd_tc_covers(const void* weight, const WFormat& f, int64_t N, int64_t K, int group32, const void* up, int64_t up_sn, int64_t up_sr,
                   const void* down, int64_t down_sr, int64_t down_sk, int rank, int svd_dtype, int out_dtype) {
    const bool rank_ok = rank == 16 || rank == 32 || rank == 64;
    const bool layout_ok = up_sr == 1 && up_sn >= rank && up_sn % 8 == 0 && down_sr == 1 && down_sk >= rank && down_sk % 8 == 0;   // both K-major
    const bool dtype_ok = svd_dtype == SDNQ_BF16 && out_dtype == SDNQ_BF16;        // bf16 model dtype (f16 operands would need kind::f16 f16 formats)
    const bool group_ok = (group32 & 7) == 0 || group32 >= K;
    const bool align_ok = (reinterpret_cast<uintptr_t>(up) & 15) == 0 && (reinterpret_cast<uintptr_t>(down) & 15) == 0 && K % 8 == 0;
    const bool fmt_ok = f.kind == SDNQ_W_INT && f.bits == 4 && K % 32 == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0;   // 16 B cp.async units
    return rank_ok && layout_ok && dtype_ok && group_ok && align_ok && fmt_ok && N * K < (int64_t(1) << 31);
}

}  // namespace

// Returns SDNQ_OK when it handled the request, 1 when the configuration is outside what the tensor-core kernel covers (the
// caller then uses the generic kernel), negative on error.
int dequant_svd_tc(const void* weight, const WFormat& f, const float* scale, const float* zp, int64_t N, int64_t K, int group32, int group_shift,
                   int gpr32, int row_stride32, const void* up, int64_t up_sn, int64_t up_sr, const void* down, int64_t down_sr, int64_t down_sk,
                   int rank, int svd_dtype, void* out, int out_dtype, cudaStream_t st) {
    if (!svd_tc_covers(weight, f, N, K, group32, up, up_sn, up_sr, down, down_sr, down_sk, rank, svd_dtype, out_dtype)) return 1;
    SvdArgs a{reinterpret_cast<const uint8_t*>(weight), scale, zp, static_cast<int>(N), static_cast<int>(K), group32, group_shift, gpr32, row_stride32, f, rank, out_dtype};
    // Narrow tiles (several CTAs per SM) for the small weights of a UNet, wider ones as the weight grows: measured on B200 over
    // the SD-XL int4 + SVD step: TN = 64 9.62 ms, 128 9.86 ms, 256 10.7 ms (SDNQ_B200_SVD_TN forces 64 / 128 / 256)
    int tn = N * K <= (int64_t(32) << 20) ? 64 : N * K <= (int64_t(128) << 20) ? 128 : 256;
    if (const char* e = getenv("SDNQ_B200_SVD_TN")) {
        const int v = atoi(e);
        if (v == 64 || v == 128 || v == 256) tn = v;
    }
    BatchEntry e;
    const int num_m = static_cast<int>((N + TM - 1) / TM);
    int rc;
    if (tn == 256) rc = fill_entry<256>(&e, a, up, up_sn, down, down_sk, out, 0);
    else if (tn == 128) rc = fill_entry<128>(&e, a, up, up_sn, down, down_sk, out, 0);
    else rc = fill_entry<64>(&e, a, up, up_sn, down, down_sk, out, 0);
    if (rc != SDNQ_OK) return rc;
    const int tiles = num_m * e.num_n;
    if (tn == 256) return launch_entries<256>(&e, nullptr, 1, tiles, rank, entry_traits(a), st);
    if (tn == 128) return launch_entries<128>(&e, nullptr, 1, tiles, rank, entry_traits(a), st);
    return launch_entries<64>(&e, nullptr, 1, tiles, rank, entry_traits(a), st);
}

// ---- batched launches: the weights of several layers dequantised by one persistent grid ------------------------------------------
size_t svd_batch_entry_bytes() { return sizeof(BatchEntry); }

// Appends the entry of one weight to a host-side table.  Returns 1 when the weight is outside what this kernel covers.
int svd_batch_fill(void* host_entry, int tn, int tile_start, const void* weight, const WFormat& f, const float* scale, const float* zp, int64_t N,
                   int64_t K, int group32, int group_shift, int gpr32, int row_stride32, const void* up, int64_t up_sn, int64_t up_sr,
                   const void* down, int64_t down_sr, int64_t down_sk, int rank, int svd_dtype, void* out, int out_dtype, int* tiles, int* traits) {
    if (!svd_tc_covers(weight, f, N, K, group32, up, up_sn, up_sr, down, down_sr, down_sk, rank, svd_dtype, out_dtype)) return 1;
    SvdArgs a{reinterpret_cast<const uint8_t*>(weight), scale, zp, static_cast<int>(N), static_cast<int>(K), group32, group_shift, gpr32, row_stride32, f, rank, out_dtype};
    BatchEntry* e = new (host_entry) BatchEntry{};
    int rc;
    if (tn == 256) rc = fill_entry<256>(e, a, up, up_sn, down, down_sk, out, tile_start);
    else if (tn == 128) rc = fill_entry<128>(e, a, up, up_sn, down, down_sk, out, tile_start);
    else rc = fill_entry<64>(e, a, up, up_sn, down, down_sk, out, tile_start);
    if (rc != SDNQ_OK) return rc;
    *tiles = static_cast<int>((N + TM - 1) / TM) * e->num_n;
    *traits |= entry_traits(a);
    return SDNQ_OK;
}

int svd_batch_run(const void* device_table, int n_entries, int total_tiles, int tn, int stage_rank, int traits, cudaStream_t st) {
    const BatchEntry* t = reinterpret_cast<const BatchEntry*>(device_table);
    if (tn == 256) return launch_entries<256>(nullptr, t, n_entries, total_tiles, stage_rank, traits, st);
    if (tn == 128) return launch_entries<128>(nullptr, t, n_entries, total_tiles, stage_rank, traits, st);
    return launch_entries<64>(nullptr, t, n_entries, total_tiles, stage_rank, traits, st);
}

}  // namespace sdnq
