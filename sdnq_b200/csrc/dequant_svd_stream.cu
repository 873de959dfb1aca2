// K3b: dequant-only path with the SVD low-rank correction as a streaming kernel over a TABLE of weights.
//
//   W[n,k] = cast( cast(q[n,k] * scale[n,g] (+ zp)) + sum_j svd_up[n,j] * svd_down[j,k] )          reference dequantizer.py:52-84
//
// One launch dequantises the weights of several layers (sdnq_b200_dequant_batch_*; prefetch.py runs it ahead of the layers'
// GEMMs), or one weight (table of one entry passed as a kernel parameter).  HBM-bound work -- 0.5 B of codes in, 2 B out per
// element -- so the design goal is the fewest instructions per element at full occupancy, not tensor throughput:
//   * a WARP owns a 128-row x 64-column tile and walks it in 16-row blocks; no shared memory, no barriers, no TMEM: every CTA
//     slot of the SM is usable and a batch keeps 16+ independent warps per SM streaming;
//   * the rank-r term of a 16 x 64 block is 8 x (r/16) mma.sync.m16n8k16 (bf16, f32 accumulate) -- 1/64 HMMA per element.  The
//     B fragments (svd_down, the tile's 64 columns) are loaded once per tile and stay in registers; the A fragments (svd_up rows)
//     are 4-byte loads from L1/L2;
//   * the n-blocks' columns are permuted so that lane (g, q) ends up with 16 CONSECUTIVE columns [16q, 16q+16) of rows g and g+8:
//     its codes are one 8-byte load per row, its results two 16-byte stores per row (4 lanes = one 128-byte line per row);
//   * rounding points as in the reference: the scaled code is rounded to bf16 (`result.to(svd dtype)`), the f32 product is added
//     and the sum rounded once (`addmm_`).
#include <cstdlib>
#include <mutex>
#include <new>

#include "hadamard_tc.cuh"     // hadtc::Half16<bf16>: cvt pack + mma.sync m16n8k16 (with their host-emulation models)
#include "unpack.cuh"

namespace sdnq {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int TN = 64;          // tile columns (one warp)

struct alignas(16) StreamEntry {
    const uint8_t* weight;      // packed 4-bit codes of the [N,K] weight
    const float* scale;
    const float* zp;
    const uint16_t* up;         // svd_up [N, r], row pitch up_pitch elements
    const uint16_t* down;       // svd_down as stored K-major: [K, r], row pitch down_pitch elements
    uint16_t* out;              // [N, K] bf16
    int N, K, rank;
    int up_pitch, down_pitch;
    int group32, group_shift, gpr32, row_stride32;
    float bias;                 // 2^23 - int_offset: turns the PRMT-built float 2^23 + code into the signed code value
    int tile_start;             // index of this weight's first tile in the launch
    int num_n;                  // tiles along K
    int tile_rows;              // rows per tile (a multiple of 16)
};

using H = hadtc::Half16<__nv_bfloat16>;

template <int KS>                // KS = k-steps of 16 of the widest rank in the launch
__global__ void __launch_bounds__(kThreads, 2)
dequant_svd_stream_kernel(const __grid_constant__ StreamEntry single, const StreamEntry* __restrict__ table, const int n_entries, const int total_tiles) {
    pdl_launch_dependents();
    pdl_wait();
    const StreamEntry* const tab = table != nullptr ? table : &single;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    int li = 0;
    for (int tile = blockIdx.x * kWarps + warp; tile < total_tiles; tile += gridDim.x * kWarps) {
        while (li + 1 < n_entries && tile >= tab[li + 1].tile_start) ++li;
        const StreamEntry e = tab[li];
        const int local = tile - e.tile_start;
        const int m0 = (local / e.num_n) * e.tile_rows, n0 = (local % e.num_n) * TN;
        const int ksteps = e.rank >> 4;
        // ---- B fragments: svd_down for the tile's 64 columns.  n-block j, fragment column nn <-> tile column 16 (nn >> 1) + 2 j + (nn & 1)
        uint32_t b[8][KS][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + 16 * (g >> 1) + 2 * j + (g & 1);
            const uint16_t* dp = e.down + int64_t(col) * e.down_pitch + 2 * q;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const bool ok = col < e.K && ks < ksteps;
                b[j][ks][0] = ok ? *reinterpret_cast<const uint32_t*>(dp + 16 * ks) : 0u;
                b[j][ks][1] = ok ? *reinterpret_cast<const uint32_t*>(dp + 16 * ks + 8) : 0u;
            }
        }
        const int c0 = n0 + 16 * q;                        // this lane's 16 columns
        const bool col_ok = c0 < e.K;                      // K % 16 == 0: all of them or none
        const int rows_end = min(e.N, m0 + e.tile_rows);
        for (int rb = m0; rb < rows_end; rb += 16) {
            const int r0 = rb + g, r1 = r0 + 8;
            const bool ok0 = r0 < rows_end, ok1 = r1 < rows_end;
            // ---- codes of both rows (8 bytes = 16 four-bit codes each), issued before the MMAs
            uint2 w0 = make_uint2(0u, 0u), w1 = make_uint2(0u, 0u);
            if (ok0 && col_ok) w0 = *reinterpret_cast<const uint2*>(e.weight + ((static_cast<uint32_t>(r0) * static_cast<uint32_t>(e.K) + c0) >> 1));
            if (ok1 && col_ok) w1 = *reinterpret_cast<const uint2*>(e.weight + ((static_cast<uint32_t>(r1) * static_cast<uint32_t>(e.K) + c0) >> 1));
            // ---- rank-r term of the 16 x 64 block
            float acc[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
            const uint16_t* u0 = e.up + int64_t(r0) * e.up_pitch + 2 * q;
            const uint16_t* u1 = e.up + int64_t(r1) * e.up_pitch + 2 * q;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                if (ks < ksteps) {
                    const uint32_t a0 = ok0 ? *reinterpret_cast<const uint32_t*>(u0 + 16 * ks) : 0u;
                    const uint32_t a1 = ok1 ? *reinterpret_cast<const uint32_t*>(u1 + 16 * ks) : 0u;
                    const uint32_t a2 = ok0 ? *reinterpret_cast<const uint32_t*>(u0 + 16 * ks + 8) : 0u;
                    const uint32_t a3 = ok1 ? *reinterpret_cast<const uint32_t*>(u1 + 16 * ks + 8) : 0u;
#pragma unroll
                    for (int j = 0; j < 8; ++j) H::mma(acc[j], a0, a1, a2, a3, b[j][ks][0], b[j][ks][1]);
                }
            }
            if (!col_ok) continue;
            // ---- dequantise + add + round + store, row r0 then row r1
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = h ? r1 : r0;
                if (!(h ? ok1 : ok0)) continue;
                const uint2 w = h ? w1 : w0;
                uint32_t o[8];
#pragma unroll
                for (int oc = 0; oc < 2; ++oc) {           // octet oc: columns c0 + 8 oc .. + 7, n-blocks j = 4 oc .. 4 oc + 3
                    const int k = c0 + 8 * oc;
                    const int grp = e.gpr32 <= 1 ? 0 : (e.group_shift >= 0 ? (k >> e.group_shift) : static_cast<int>(static_cast<uint32_t>(k) / static_cast<uint32_t>(e.group32)));
                    const uint32_t si = static_cast<uint32_t>(r) * static_cast<uint32_t>(e.row_stride32) + grp;
                    const float sc = e.scale[si];
                    const float z = e.zp != nullptr ? e.zp[si] : 0.f;
                    const uint32_t word = oc ? w.y : w.x;
                    const uint32_t lo = word & 0x0F0F0F0Fu, hi = (word >> 4) & 0x0F0F0F0Fu;      // even / odd elements of the octet
#pragma unroll
                    for (int p = 0; p < 4; ++p) {          // byte p = elements (2p, 2p+1) = n-block j = 4 oc + p, fragment columns (2q, 2q+1)
                        const float q0 = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7440 | p)) - e.bias;
                        const float q1 = __uint_as_float(__byte_perm(hi, 0x4B000000u, 0x7440 | p)) - e.bias;
                        const float v0 = e.zp != nullptr ? fmaf(q0, sc, z) : __fmul_rn(q0, sc);
                        const float v1 = e.zp != nullptr ? fmaf(q1, sc, z) : __fmul_rn(q1, sc);
                        const uint32_t wb = H::pack(v0, v1);                                       // result.to(svd dtype): bf16
                        const int j = 4 * oc + p;
                        o[j] = H::pack(__fadd_rn(H::lo(wb), acc[j][2 * h]), __fadd_rn(H::hi(wb), acc[j][2 * h + 1]));   // addmm_: f32 sum, rounded once
                    }
                }
                uint4* dst = reinterpret_cast<uint4*>(e.out + (static_cast<uint32_t>(r) * static_cast<uint32_t>(e.K) + c0));
                dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
    }
}

template <int KS>
int launch_stream(const StreamEntry* single, const StreamEntry* device_table, int n_entries, int total_tiles, cudaStream_t st) {
    // two CTAs of 8 warps per SM; a launch with fewer tiles than that gets one warp per tile
    const int64_t ctas = (int64_t(total_tiles) + kWarps - 1) / kWarps;
    const int64_t cap = int64_t(num_sms()) * 2;
    const unsigned grid = static_cast<unsigned>(ctas < cap ? ctas : cap);
    static const StreamEntry kNone{};
    cudaError_t e = launch_pdl(dequant_svd_stream_kernel<KS>, dim3(grid), dim3(kThreads), 0, st, single != nullptr ? *single : kNone, device_table,
                               n_entries, total_tiles);
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of dequant_svd_stream_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("dequant_svd_stream_kernel");
}

int launch_stream_rank(int max_rank, const StreamEntry* single, const StreamEntry* device_table, int n_entries, int total_tiles, cudaStream_t st) {
    if (max_rank <= 16) return launch_stream<1>(single, device_table, n_entries, total_tiles, st);
    if (max_rank <= 32) return launch_stream<2>(single, device_table, n_entries, total_tiles, st);
    return launch_stream<4>(single, device_table, n_entries, total_tiles, st);
}

bool stream_covers(const void* weight, const WFormat& f, int64_t N, int64_t K, int group32, const void* up, int64_t up_sn, int64_t up_sr,
                   const void* down, int64_t down_sr, int64_t down_sk, int rank, int svd_dtype, const void* out, int out_dtype) {
    const bool rank_ok = rank == 16 || rank == 32 || rank == 48 || rank == 64;
    const bool layout_ok = up_sr == 1 && up_sn >= rank && up_sn % 2 == 0 && down_sr == 1 && down_sk >= rank && down_sk % 2 == 0;   // both K-major
    const bool dtype_ok = svd_dtype == SDNQ_BF16 && out_dtype == SDNQ_BF16;
    const bool group_ok = (group32 & 7) == 0 || group32 >= K;
    const bool align_ok = (reinterpret_cast<uintptr_t>(up) & 3) == 0 && (reinterpret_cast<uintptr_t>(down) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const bool fmt_ok = f.kind == SDNQ_W_INT && f.bits == 4 && f.word_bytes == 1 && K % 16 == 0 && (reinterpret_cast<uintptr_t>(weight) & 7) == 0;
    return rank_ok && layout_ok && dtype_ok && group_ok && align_ok && fmt_ok && N * K < (int64_t(1) << 31) && up_sn < (int64_t(1) << 31) && down_sk < (int64_t(1) << 31);
}

void fill_stream(StreamEntry* e, const void* weight, const WFormat& f, const float* scale, const float* zp, int64_t N, int64_t K, int group32, int group_shift,
                 int gpr32, int row_stride32, const void* up, int64_t up_sn, const void* down, int64_t down_sk, int rank, void* out, int tile_start, int tile_rows) {
    e->weight = reinterpret_cast<const uint8_t*>(weight);
    e->scale = scale;
    e->zp = zp;
    e->up = reinterpret_cast<const uint16_t*>(up);
    e->down = reinterpret_cast<const uint16_t*>(down);
    e->out = reinterpret_cast<uint16_t*>(out);
    e->N = static_cast<int>(N);
    e->K = static_cast<int>(K);
    e->rank = rank;
    e->up_pitch = static_cast<int>(up_sn);
    e->down_pitch = static_cast<int>(down_sk);
    e->group32 = group32;
    e->group_shift = group_shift;
    e->gpr32 = gpr32;
    e->row_stride32 = row_stride32;
    e->bias = 8388608.0f - static_cast<float>(f.int_offset);
    e->tile_start = tile_start;
    e->num_n = static_cast<int>((K + TN - 1) / TN);
    e->tile_rows = tile_rows;
}

int tiles_of(const StreamEntry& e) { return ((e.N + e.tile_rows - 1) / e.tile_rows) * e.num_n; }

}  // namespace

// One weight.  Returns SDNQ_OK when it handled the request, 1 when the configuration is outside what the kernel covers.
int dequant_svd_stream(const void* weight, const WFormat& f, const float* scale, const float* zp, int64_t N, int64_t K, int group32, int group_shift,
                       int gpr32, int row_stride32, const void* up, int64_t up_sn, int64_t up_sr, const void* down, int64_t down_sr, int64_t down_sk,
                       int rank, int svd_dtype, void* out, int out_dtype, cudaStream_t st) {
    if (!stream_covers(weight, f, N, K, group32, up, up_sn, up_sr, down, down_sr, down_sk, rank, svd_dtype, out, out_dtype)) return 1;
    // rows per tile: 128 for a weight with plenty of tiles, fewer rows (more warps) for the small weights of a UNet
    const int64_t strips = (K + TN - 1) / TN;
    const int64_t want = int64_t(num_sms()) * 2 * kWarps * 2;          // two tiles per warp slot
    int tile_rows = 128;
    while (tile_rows > 16 && ((N + tile_rows - 1) / tile_rows) * strips < want) tile_rows >>= 1;
    StreamEntry e;
    fill_stream(&e, weight, f, scale, zp, N, K, group32, group_shift, gpr32, row_stride32, up, up_sn, down, down_sk, rank, out, 0, tile_rows);
    return launch_stream_rank(rank, &e, nullptr, 1, tiles_of(e), st);
}

// ---- batched launches: the weights of several layers dequantised by one grid ----------------------------------------------------
size_t svd_batch_entry_bytes() { return sizeof(StreamEntry); }

// Appends the entry of one weight to a host-side table.  Returns 1 when the weight is outside what this kernel covers.
int svd_batch_fill(void* host_entry, int tile_rows, int tile_start, const void* weight, const WFormat& f, const float* scale, const float* zp, int64_t N,
                   int64_t K, int group32, int group_shift, int gpr32, int row_stride32, const void* up, int64_t up_sn, int64_t up_sr,
                   const void* down, int64_t down_sr, int64_t down_sk, int rank, int svd_dtype, void* out, int out_dtype, int* tiles) {
    if (!stream_covers(weight, f, N, K, group32, up, up_sn, up_sr, down, down_sr, down_sk, rank, svd_dtype, out, out_dtype)) return 1;
    StreamEntry* e = new (host_entry) StreamEntry{};
    fill_stream(e, weight, f, scale, zp, N, K, group32, group_shift, gpr32, row_stride32, up, up_sn, down, down_sk, rank, out, tile_start, tile_rows);
    *tiles = tiles_of(*e);
    return SDNQ_OK;
}

int svd_batch_run(const void* device_table, int n_entries, int total_tiles, int max_rank, cudaStream_t st) {
    return launch_stream_rank(max_rank, nullptr, reinterpret_cast<const StreamEntry*>(device_table), n_entries, total_tiles, st);
}

}  // namespace sdnq
