// K3 (dequant-only HBM kernel), the unpack-only kernel and K4 (re-quantise for matmul).
//
// Reference behaviour restated here:
//   dequantize_symmetric / _asymmetric / _codebook   dequantizer.py:15-131
//   dequantize_weight / SDNQDequantizer.__call__      dequantizer.py:135-162, 389-429
//   re_quantize_{int,uint,fp}_mm / re_quantize_matmul dequantizer.py:166-239
//   quantize_{int,uint,fp}_mm                         quant_utils.py:264-299
//
// Data movement: one lane owns one *octet* (8 consecutive values of a row = `bits` storage bytes), so a warp
// reads 32*bits contiguous bytes and writes 256 contiguous outputs (512 B of bf16) per step -- every global
// access is a full-sector, fully-coalesced transaction and there is no reuse to stage in shared memory.
#include "unpack.cuh"
#include "hadamard_tc.cuh"

namespace sdnq {

size_t svd_batch_entry_bytes();
int svd_batch_fill(void* host_entry, int tn, int tile_start, const void* weight, const WFormat& f, const float* scale, const float* zp, int64_t N,
                   int64_t K, int group32, int group_shift, int gpr32, int row_stride32, const void* up, int64_t up_sn, int64_t up_sr,
                   const void* down, int64_t down_sr, int64_t down_sk, int rank, int svd_dtype, void* out, int out_dtype, int* tiles, int* traits);
int svd_batch_run(const void* device_table, int n_entries, int total_tiles, int tn, int stage_rank, int traits, cudaStream_t st);
int dequant_svd_tc(const void* weight, const WFormat& f, const float* scale, const float* zp, int64_t N, int64_t K, int group32, int group_shift,
                   int gpr32, int row_stride32, const void* up, int64_t up_sn, int64_t up_sr, const void* down, int64_t down_sr, int64_t down_sk,
                   int rank, int svd_dtype, void* out, int out_dtype, cudaStream_t st);

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct DequantArgs {
    const uint8_t* weight;
    const float* scale;
    const float* zp;
    int64_t N, K;
    int64_t group;          // elements per scale along K (K for row-wise); -2 tensor-wise handled as group = K with stride 0
    int64_t groups_per_row; // K / group (0 stride for tensor-wise)
    int64_t scale_row_stride;
    int codebook;           // scale holds 2^bits levels per group
    WFormat f;
    // svd
    const void* up; int64_t up_sn, up_sr;
    const void* down; int64_t down_sr, down_sk;
    int rank; int svd_dtype;
    int hadamard;
    // 32-bit copies for the hot kernel (N*K < 2^31 elements is enforced on the host)
    int K32, group32, group_shift;   // group_shift >= 0 when the group size is a power of two
    int gpr32, row_stride32;
    // embedding lookup (quantized_embedding, layers/embedding/forward.py:14-68): output row n is stored row gather[n] of `src_rows`
    const int64_t* gather;
    int64_t src_rows;
    float out_scale;        // embed_scale: result.mul_(embed_scale) in the result dtype (1 = none)
};

__device__ __forceinline__ float load_any(const void* p, int64_t i, int dtype) {
    if (dtype == SDNQ_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    if (dtype == SDNQ_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
    return reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ float round_any(float v, int dtype) {
    if (dtype == SDNQ_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    if (dtype == SDNQ_F16) return __half2float(__float2half_rn(v));
    return v;
}

__device__ __forceinline__ int group_of(const DequantArgs& a, int k) {
    if (a.gpr32 <= 1) return 0;
    return a.group_shift >= 0 ? (k >> a.group_shift) : static_cast<int>(static_cast<uint32_t>(k) / static_cast<uint32_t>(a.group32));
}

// scale (+ zero point | codebook) of the octet starting at (row n, column k): f32 values, reference op order
__device__ __forceinline__ void scale_octet(const DequantArgs& a, int n, int k, const float (&q)[8], const uint32_t (&codes)[8],
                                            float (&w)[8]) {
    if ((a.group32 & 7) == 0 || a.group32 >= a.K32) {
        const int64_t si = int64_t(n) * a.row_stride32 + group_of(a, k);
        if (a.codebook) {
            const float* lv = a.scale + si * (int64_t(1) << a.f.bits);
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = lv[codes[i]];
        } else {
            const float s = a.scale[si];
            if (a.zp != nullptr) {
                const float z = a.zp[si];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = fmaf(q[i], s, z);      // addcmul(zp, q, scale): one fused op in ATen
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = __fmul_rn(q[i], s);
            }
        }
    } else {   // group sizes that are not a multiple of 8: per-element scale lookup
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t si = int64_t(n) * a.row_stride32 + group_of(a, k + i);
            if (a.codebook) w[i] = a.scale[si * (int64_t(1) << a.f.bits) + codes[i]];
            else if (a.zp != nullptr) w[i] = fmaf(q[i], a.scale[si], a.zp[si]);
            else w[i] = __fmul_rn(q[i], a.scale[si]);
        }
    }
}

// unpack + scale of one octet (used by the re-quantise kernel)
template <int BITS>
__device__ __forceinline__ void dequant_octet(const DequantArgs& a, int64_t n, int64_t k, float (&w)[8]) {
    uint32_t codes[8];
    float q[8];
    octet_values<BITS>(a.weight, (n * a.K + k) >> 3, a.f, q, codes);
    scale_octet(a, static_cast<int>(n), static_cast<int>(k), q, codes, w);
}

// ------------------------------------------------------------------------------------------------ K3
// Work unit = one warp x one 256-element chunk of one row (lane l owns the octet at column 256*c + 8*l).  A warp walks
// chunks with a grid stride and keeps U chunks in flight: all U packed-byte loads are issued before any of them is used,
// which is what the kernel needs to cover HBM latency (it has no reuse, so memory-level parallelism is the only lever:
// ~2k threads/SM x U independent loads).  Reads: 32*bits contiguous bytes per warp per chunk; writes: 512 contiguous
// bytes of bf16 per warp per chunk (one 16 B store per lane).
template <int BITS, typename OutT, int U, bool kPlain>
__global__ void __launch_bounds__(kThreads, kPlain ? 5 : 3) dequant_kernel(const DequantArgs a, OutT* __restrict__ out, int chunks_per_row, int total_chunks) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * kWarps;
    for (int t0 = warp_global; t0 < total_chunks; t0 += warps_total * U) {
        int n[U], k[U], ns[U];
        bool live[U], valid[U];
        uint32_t raw[U][OctetWords<BITS>::N];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = t0 + u * warps_total;
            live[u] = t < total_chunks;                               // warp-uniform
            n[u] = live[u] ? static_cast<int>(static_cast<uint32_t>(t) / static_cast<uint32_t>(chunks_per_row)) : 0;
            k[u] = live[u] ? (t - n[u] * chunks_per_row) * 256 + lane * 8 : 0;
            valid[u] = live[u] && k[u] < a.K32;
            ns[u] = n[u];
            if (!kPlain && a.gather != nullptr && live[u]) {          // embedding lookup: the stored row this output row comes from
                const int64_t g = a.gather[n[u]];
                if (g < 0 || g >= a.src_rows) __trap();               // an index outside the table (torch raises a device-side assert here too)
                ns[u] = static_cast<int>(g);
            }
            if (valid[u]) load_octet_bytes<BITS>(a.weight, (int64_t(ns[u]) * a.K32 + k[u]) >> 3, a.f.word_bytes, raw[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!live[u]) continue;
            float w[8];
            if (valid[u]) {
                uint32_t codes[8];
                float q[8];
                decode_octet<BITS>(raw[u], codes);
                codes_to_values<BITS>(codes, a.f, q);
                scale_octet(a, ns[u], k[u], q, codes, w);
                if constexpr (!kPlain) {
                    if (a.up != nullptr) {
                        // result.to(svd dtype).addmm_(svd_up, svd_down): f32 accumulate, one rounding (dequantizer.py:69-79)
                        float acc[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] = round_any(w[i], a.svd_dtype);
                        for (int j = 0; j < a.rank; ++j) {
                            const float uj = load_any(a.up, int64_t(ns[u]) * a.up_sn + j * a.up_sr, a.svd_dtype);
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                acc[i] = fmaf(uj, load_any(a.down, j * a.down_sr + int64_t(k[u] + i) * a.down_sk, a.svd_dtype), acc[i]);
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) w[i] = round_any(acc[i], a.svd_dtype);
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = ElemTraits<OutT>::round(w[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = 0.f;
            }
            if constexpr (!kPlain) {
                if (a.hadamard) {   // un-rotate in the result dtype (dequantizer.py:82-83); whole warp participates
                    hadamard_warp_dyn(a.hadamard, w, hadamard_factor<OutT>(a.hadamard));
                }
                if (a.out_scale != 1.f) {                                 // embedding: result.mul_(embed_scale) on the rounded values
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = ElemTraits<OutT>::round(__fmul_rn(ElemTraits<OutT>::round(w[i]), a.out_scale));
                }
            }
            if (valid[u]) {
                if (kPlain || !a.hadamard) {
                    store8<OutT>(out + int64_t(n[u]) * a.K32 + k[u], w);
                } else {   // power-of-4 un-rotation leaves the two halves of the lane at permuted places of the chunk
                    OutT* o = out + int64_t(n[u]) * a.K32 + (k[u] - lane * 8);
                    store4<OutT>(o + hadamard_dest_dyn(a.hadamard, lane, 0), w[0], w[1], w[2], w[3]);
                    store4<OutT>(o + hadamard_dest_dyn(a.hadamard, lane, 1), w[4], w[5], w[6], w[7]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3 rotated 8-bit path
// 8-bit formats (int8 / uint8 / fp8 / 8-bit minifloats) whose dequantised weight is un-rotated (use_hadamard layers on the
// dequant path: FLUX AdaLN Linears with M = batch < 32, `.dequantize()`): scale -> round to the result dtype -> Hadamard on
// the tensor cores (hadamard_tc.cuh) -> store.  Lane l owns elements [4l, 4l+4) and [128+4l, 128+4l+4) of a 256-chunk: two
// 4-byte code loads, two 8-byte stores, all coalesced.  The generic kernel's shuffle butterflies made this path issue-bound
// (1.7 TB/s on the 3072 -> 18432 AdaLN weight).
template <typename OutT, int U>
__global__ void __launch_bounds__(kThreads, 4) dequant_rot8_kernel(const DequantArgs a, OutT* __restrict__ out, int chunks_per_row, int total_chunks) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * kWarps;
    hadtc::Rotation<OutT> rot;
    rot.init(a.hadamard, lane);
    const float factor = hadamard_factor<OutT>(a.hadamard);
    for (int t0 = warp_global; t0 < total_chunks; t0 += warps_total * U) {
        int n[U], k[U];
        bool live[U];
        uint32_t lo[U], hi[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = t0 + u * warps_total;
            live[u] = t < total_chunks;                               // warp-uniform
            n[u] = live[u] ? static_cast<int>(static_cast<uint32_t>(t) / static_cast<uint32_t>(chunks_per_row)) : 0;
            k[u] = live[u] ? (t - n[u] * chunks_per_row) * 256 + lane * 4 : 0;
            lo[u] = hi[u] = 0u;
            const uint8_t* row = a.weight + int64_t(n[u]) * a.K32;
            if (live[u] && k[u] < a.K32) lo[u] = *reinterpret_cast<const uint32_t*>(row + k[u]);
            if (live[u] && k[u] + 128 < a.K32) hi[u] = *reinterpret_cast<const uint32_t*>(row + k[u] + 128);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!live[u]) continue;
            const bool ok_lo = k[u] < a.K32, ok_hi = k[u] + 128 < a.K32;
            uint32_t codes[8];
            float q[8], w[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) { codes[i] = (lo[u] >> (8 * i)) & 0xFFu; codes[4 + i] = (hi[u] >> (8 * i)) & 0xFFu; }
            codes_to_values<8>(codes, a.f, q);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool ok = h ? ok_hi : ok_lo;
                const int64_t si = int64_t(n[u]) * a.row_stride32 + group_of(a, k[u] + 128 * h);
                const float sc = ok ? a.scale[si] : 0.f;
                const float z = (ok && a.zp != nullptr) ? a.zp[si] : 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float v = a.zp != nullptr ? fmaf(q[4 * h + i], sc, z) : __fmul_rn(q[4 * h + i], sc);
                    w[4 * h + i] = ok ? v : 0.f;
                }
            }
            uint4 raw;                                                // rounded to the result dtype before the rotation (dequantizer.py:80-83)
            raw.x = hadtc::Half16<OutT>::pack(w[0], w[1]);
            raw.y = hadtc::Half16<OutT>::pack(w[2], w[3]);
            raw.z = hadtc::Half16<OutT>::pack(w[4], w[5]);
            raw.w = hadtc::Half16<OutT>::pack(w[6], w[7]);
            rot.apply(raw, factor);
            OutT* o = out + int64_t(n[u]) * a.K32 + k[u];
            if (ok_lo) *reinterpret_cast<uint2*>(o) = make_uint2(raw.x, raw.y);
            if (ok_hi) *reinterpret_cast<uint2*>(o + 128) = make_uint2(raw.z, raw.w);
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3 fast path
// Integer formats (int/uint 1..8 bit), one scale (and zero point) per octet, no codebook / SVD / Hadamard: everything that can
// be decided at compile time is.  Each warp streams a contiguous run of 256-element chunks (row / column advance
// incrementally: one integer division per warp, none per chunk) with U chunks in flight; int4 and int8 go from storage
// word to float with one PRMT per element (byte -> 0x4B0000bb = 2^23 + b, then one FADD), so the kernel sits at ~7 SASS
// instructions per element and 40 registers (6 CTAs / SM).
template <int BITS, bool kZP, bool kPow2, typename OutT, int U>
__global__ void __launch_bounds__(kThreads, 6) dequant_int_kernel(const DequantArgs a, OutT* __restrict__ out, int cpr, int total_chunks) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * kWarps;
    const bool twos = (BITS == 8 && !a.f.is_unsigned);
    const float bias = twos ? 8388736.0f : 8388608.0f - static_cast<float>(a.f.int_offset);
    const uint32_t flip = twos ? 0x80808080u : 0u;
    const int per_warp = (total_chunks + warps_total - 1) / warps_total;
    int t = warp_global * per_warp;
    const int t_end = min(t + per_warp, total_chunks);
    if (t >= t_end) return;
    int n = static_cast<int>(static_cast<uint32_t>(t) / static_cast<uint32_t>(cpr));
    int c = t - n * cpr;
    const int K = a.K32, wb = a.f.word_bytes;
    for (; t < t_end; t += U) {
        int nn[U], kk[U];
        bool valid[U];
        uint32_t raw[U][OctetWords<BITS>::N];
        float sc[U], zp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            nn[u] = n;
            kk[u] = c * 256 + lane * 8;
            valid[u] = (t + u < t_end) && kk[u] < K;
            if (++c == cpr) { c = 0; ++n; }
            sc[u] = zp[u] = 0.f;
            if (valid[u]) {
                const uint32_t e = static_cast<uint32_t>(nn[u]) * static_cast<uint32_t>(K) + static_cast<uint32_t>(kk[u]);   // < 2^31
                load_octet_bytes<BITS>(a.weight, e >> 3, wb, raw[u]);
                const int g = a.gpr32 <= 1 ? 0 : (kPow2 ? (kk[u] >> a.group_shift) : static_cast<int>(static_cast<uint32_t>(kk[u]) / static_cast<uint32_t>(a.group32)));
                const uint32_t si = static_cast<uint32_t>(nn[u]) * static_cast<uint32_t>(a.row_stride32) + g;
                sc[u] = a.scale[si];
                if constexpr (kZP) zp[u] = a.zp[si];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!valid[u]) continue;
            float q[8], w[8];
            octet_to_floats<BITS>(raw[u], flip, bias, q);
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = kZP ? fmaf(q[i], sc[u], zp[u]) : __fmul_rn(q[i], sc[u]);
            store8<OutT>(out + (static_cast<uint32_t>(nn[u]) * static_cast<uint32_t>(K) + static_cast<uint32_t>(kk[u])), w);
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3 flat path (int4 / int8)
// The weight is one contiguous [N*K] array and, when every scale group lies inside a row and the scale rows are dense, the
// scale of flat element e is simply scale[e / group]: no row / column arithmetic at all.  A CTA takes U blocks of 256
// consecutive octets; thread t owns octet (block*256 + t) of each, so every load (4 or 8 B per lane) and every store
// (16 B per lane, 512 B per warp) is a fully coalesced warp access.  Per octet: 1-2 LDG, 8 x (PRMT, FADD, FMUL | FFMA),
// 4 pack, 1 STG.128 -- ~4.5 SASS instructions per element, a third of the row-walking kernel, which leaves the
// kernel bound by HBM alone.  kRow: row-wise scale (group == K): the row of an octet by multiply-high division.
// kFp8: the bytes are float8_e4m3fn values (cvt.f16x2.e4m3x2, exact) instead of integer codes.
template <int BITS, bool kZP, bool kRow, typename OutT, int U, bool kFp8 = false>
__global__ void __launch_bounds__(kThreads) dequant_flat_kernel(const uint8_t* __restrict__ weight, const float* __restrict__ scale,
                                                                const float* __restrict__ zp, OutT* __restrict__ out,
                                                                uint32_t total_octets, int shift, uint32_t opr, uint32_t opr_magic,
                                                                uint32_t flip, float bias, int streaming) {
    static_assert(BITS == 4 || BITS == 8, "flat path: one or two storage words per octet");
    pdl_launch_dependents();
    pdl_wait();
    constexpr int W = BITS / 4;
    const uint32_t stride = gridDim.x * uint32_t(kThreads * U);
    for (uint32_t base = blockIdx.x * uint32_t(kThreads * U) + threadIdx.x; base < total_octets; base += stride) {
        uint32_t raw[U][W];
        float sc[U], z[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t o = base + uint32_t(u * kThreads);
            sc[u] = z[u] = 0.f;
            if (o < total_octets) {
                if constexpr (W == 1) {
                    raw[u][0] = reinterpret_cast<const uint32_t*>(weight)[o];
                } else {
                    const uint2 v = reinterpret_cast<const uint2*>(weight)[o];
                    raw[u][0] = v.x;
                    raw[u][1] = v.y;
                }
                uint32_t si;
                if constexpr (kRow) {
                    si = __umulhi(o, opr_magic);                       // floor(o / opr) or one less
                    if (o - si * opr >= opr) ++si;
                } else {
                    si = o >> shift;
                }
                sc[u] = scale[si];
                if constexpr (kZP) z[u] = zp[si];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t o = base + uint32_t(u * kThreads);
            if (o >= total_octets) continue;
            float q[8], w[8];
            if constexpr (kFp8) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t h;
#ifdef SDNQ_HOST_EMU
                    h = ::sdnq_emu::e4m3x2_to_f16x2(static_cast<unsigned short>((raw[u][i >> 1] >> (16 * (i & 1))) & 0xFFFFu));
#else
                    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h) : "h"(static_cast<unsigned short>((raw[u][i >> 1] >> (16 * (i & 1))) & 0xFFFFu)));
#endif
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
                    q[2 * i] = f.x;
                    q[2 * i + 1] = f.y;
                }
            } else {
                octet_to_floats<BITS>(raw[u], flip, bias, q);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = kZP ? fmaf(q[i], sc[u], z[u]) : __fmul_rn(q[i], sc[u]);
            if constexpr (sizeof(OutT) == 2) {
                if (streaming) {            // the dequantised weight is consumed once by the GEMM behind us: evict-first stores
                    uint32_t pk[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if constexpr (ElemTraits<OutT>::kDtype == SDNQ_BF16) {
                            __nv_bfloat162 h = __floats2bfloat162_rn(w[2 * i], w[2 * i + 1]);
                            pk[i] = *reinterpret_cast<uint32_t*>(&h);
                        } else {
                            __half2 h = __floats2half2_rn(w[2 * i], w[2 * i + 1]);
                            pk[i] = *reinterpret_cast<uint32_t*>(&h);
                        }
                    }
#ifdef SDNQ_HOST_EMU
                    *reinterpret_cast<uint4*>(out + size_t(o) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
#else
                    asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(out + size_t(o) * 8), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
#endif
                    continue;
                }
            }
            store8<OutT>(out + size_t(o) * 8, w);
        }
    }
}

// ------------------------------------------------------------------------------------------------ unpack only
template <int BITS>
__global__ void __launch_bounds__(kThreads) unpack_kernel(const uint8_t* __restrict__ packed, WFormat f, void* __restrict__ out,
                                                          int out_dtype, int64_t octets) {
    // no early trigger: the codes written here may be the B operand of the GEMM, which prefetches its weights *before* its own
    // griddepcontrol.wait (weights are assumed complete when a dependent starts); dependents start when this grid has finished
    pdl_wait();
    const int64_t oct = int64_t(blockIdx.x) * kThreads + threadIdx.x;
    if (oct >= octets) return;
    uint32_t codes[8];
    float q[8];
    octet_values<BITS>(packed, oct, f, q, codes);
    const int64_t e = oct * 8;
    if (out_dtype == SDNQ_I8 || out_dtype == SDNQ_U8) {
        uint2 r;
        uint8_t* b = reinterpret_cast<uint8_t*>(&r);
#pragma unroll
        for (int i = 0; i < 8; ++i) b[i] = static_cast<uint8_t>(static_cast<int>(q[i]));
        *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(out) + e) = r;
    } else if (out_dtype == SDNQ_F8E4M3) {
        uint2 r;
        uint8_t* b = reinterpret_cast<uint8_t*>(&r);
#pragma unroll
        for (int i = 0; i < 8; ++i) b[i] = f32_to_e4m3(fminf(fmaxf(q[i], -448.f), 448.f));
        *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(out) + e) = r;
    } else if (out_dtype == SDNQ_I32) {
        int* o = reinterpret_cast<int*>(out) + e;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = static_cast<int>(q[i]);
    } else if (out_dtype == SDNQ_BF16) {
        store8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(out) + e, q);
    } else if (out_dtype == SDNQ_F16) {
        store8<__half>(reinterpret_cast<__half*>(out) + e, q);
    } else {
        store8<float>(reinterpret_cast<float*>(out) + e, q);
    }
}

// ------------------------------------------------------------------------------------------------ K4
// One CTA per output row n: dequant the row to f32 registers, block-reduce amax (or min/max), re-quantise.
constexpr int kRequantMaxOct = 8;   // octets per thread: K <= 8 * 8 * 256 = 16384

__device__ __forceinline__ float block_reduce_max(float v, float* s) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = s[0];
#pragma unroll
    for (int i = 1; i < kWarps; ++i) r = fmaxf(r, s[i]);
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_reduce_min(float v, float* s) {
    v = warp_min(v);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = s[0];
#pragma unroll
    for (int i = 1; i < kWarps; ++i) r = fminf(r, s[i]);
    __syncthreads();
    return r;
}

template <int BITS>
__global__ void __launch_bounds__(kThreads) requant_kernel(const DequantArgs a, int mm_dtype, uint8_t* __restrict__ wq,
                                                           float* __restrict__ sw, float* __restrict__ zw,
                                                           int32_t* __restrict__ colsum) {
    __shared__ float s_red[kWarps];
    __shared__ int s_sum[kWarps];
    // no early trigger (see unpack_kernel): wq / sw written here are the GEMM's weight operand
    pdl_wait();
    const int64_t n = blockIdx.x;
    float w[kRequantMaxOct][8];
    float vmax = -INFINITY, vmin = INFINITY, amax = 0.f;
#pragma unroll
    for (int o = 0; o < kRequantMaxOct; ++o) {
        const int64_t k = (int64_t(o) * kThreads + threadIdx.x) * 8;
        if (k < a.K) {
            dequant_octet<BITS>(a, n, k, w[o]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                amax = fmaxf(amax, fabsf(w[o][i]));
                vmax = fmaxf(vmax, w[o][i]);
                vmin = fminf(vmin, w[o][i]);
            }
        }
    }
    float scale, zero = 0.f;
    if (mm_dtype == SDNQ_U8) {          // quantize_uint_mm -> get_scale_asymmetric(.., "int8")  quant_utils.py:9-19, 276-286
        vmax = block_reduce_max(vmax, s_red);
        vmin = block_reduce_min(vmin, s_red);
        scale = __fdiv_rn(__fsub_rn(vmax, vmin), 255.f);
        zero = __fsub_rn(vmin, __fmul_rn(scale, -128.f));
    } else {                            // get_scale_symmetric                                      quant_utils.py:22-24
        amax = block_reduce_max(amax, s_red);
        scale = __fdiv_rn(amax, mm_dtype == SDNQ_F8E4M3 ? 448.f : 127.f);
    }
    int local_sum = 0;
#pragma unroll
    for (int o = 0; o < kRequantMaxOct; ++o) {
        const int64_t k = (int64_t(o) * kThreads + threadIdx.x) * 8;
        if (k < a.K) {
            uint2 r;
            uint8_t* b = reinterpret_cast<uint8_t*>(&r);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = w[o][i];
                if (mm_dtype == SDNQ_U8) v = __fsub_rn(v, zero);
                v = __fdiv_rn(v, scale);
                if (mm_dtype == SDNQ_F8E4M3) {
                    if (v != v) v = 0.f;                                        // nan_to_num
                    v = fminf(fmaxf(v, -448.f), 448.f);
                    b[i] = f32_to_e4m3(v);
                } else {
                    v = rintf(v);
                    const int c = (v != v) ? 0 : static_cast<int>(fminf(fmaxf(v, -128.f), 127.f));
                    b[i] = static_cast<uint8_t>(static_cast<int8_t>(c));
                    local_sum += c;
                }
            }
            *reinterpret_cast<uint2*>(wq + n * a.K + k) = r;
        }
    }
    if (colsum != nullptr) {
        local_sum = warp_sum(local_sum);
        if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = local_sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
#pragma unroll
            for (int i = 0; i < kWarps; ++i) t += s_sum[i];
            colsum[n] = t;
        }
    }
    if (threadIdx.x == 0) {
        sw[n] = scale;
        if (zw != nullptr) zw[n] = zero;
    }
}

int fill_args(DequantArgs& a, const void* weight, const sdnq_weight_format* fmt, const float* scale, const float* zp,
              int codebook, int64_t N, int64_t K, int64_t group_size) {
    int rc = make_wformat(fmt, &a.f);
    if (rc != SDNQ_OK) return rc;
    SDNQ_REQUIRE(weight != nullptr && scale != nullptr, SDNQ_EINVAL, "weight/scale pointer is NULL");
    SDNQ_REQUIRE(N > 0 && K > 0, SDNQ_EINVAL, "empty weight: N=%lld K=%lld", (long long)N, (long long)K);
    SDNQ_REQUIRE(K % 8 == 0, SDNQ_EUNSUPPORTED, "K (=%lld) must be a multiple of 8", (long long)K);
    SDNQ_REQUIRE((reinterpret_cast<uintptr_t>(weight) & 7) == 0, SDNQ_EINVAL, "weight pointer must be 8-byte aligned");
    SDNQ_REQUIRE(N <= 0x7fffffffLL && K <= 65535LL * 4096, SDNQ_EUNSUPPORTED, "weight too large: N=%lld K=%lld", (long long)N, (long long)K);
    a.weight = reinterpret_cast<const uint8_t*>(weight);
    a.scale = scale;
    a.zp = zp;
    a.N = N;
    a.K = K;
    a.codebook = codebook;
    if (group_size == -2) {              // tensor-wise: one scalar
        a.group = K; a.groups_per_row = 0; a.scale_row_stride = 0;
    } else {
        if (group_size <= 0 || group_size > K) group_size = K;
        SDNQ_REQUIRE(K % group_size == 0, SDNQ_EINVAL, "group_size %lld does not divide K %lld", (long long)group_size, (long long)K);
        a.group = group_size; a.groups_per_row = K / group_size; a.scale_row_stride = K / group_size;
    }
    if (codebook) SDNQ_REQUIRE(a.f.kind == SDNQ_W_INT && a.f.is_unsigned, SDNQ_EINVAL, "codebook needs an unsigned integer format");
    SDNQ_REQUIRE(N * K < (int64_t(1) << 31), SDNQ_EUNSUPPORTED, "weights with 2^31 or more elements are not supported (N=%lld K=%lld)", (long long)N, (long long)K);
    a.K32 = static_cast<int>(K);
    a.group32 = static_cast<int>(a.group);
    a.group_shift = -1;
    if ((a.group32 & (a.group32 - 1)) == 0) {
        a.group_shift = 0;
        while ((1 << a.group_shift) < a.group32) ++a.group_shift;
    }
    a.gpr32 = static_cast<int>(a.groups_per_row);
    a.row_stride32 = static_cast<int>(a.scale_row_stride);
    a.up = a.down = nullptr;
    a.up_sn = a.up_sr = a.down_sr = a.down_sk = 0;
    a.rank = 0; a.svd_dtype = SDNQ_BF16; a.hadamard = 0;
    a.gather = nullptr; a.src_rows = N; a.out_scale = 1.f;
    return SDNQ_OK;
}

// Ordering barrier behind the kernels that *produce a GEMM weight operand* (unpack, re-quantise).  The GEMM prefetches its
// weight tiles before its griddepcontrol.wait (weights are assumed complete and visible when a programmatic dependent
// starts), and only a wait / a normally launched kernel gives that guarantee: this empty kernel is launched without the
// programmatic attribute, so it starts after the producer has completed and flushed, and it never triggers early, so whatever
// follows starts after it.  Runs once per layer (the operands are cached).
__global__ void operand_fence_kernel() {}

bool hadamard_ok(int g) { return g == 0 || (g >= 4 && g <= 256 && (g & (g - 1)) == 0); }

}  // namespace

template <typename OutT>
static int launch_dequant(const DequantArgs& a, void* out, cudaStream_t st) {
    constexpr int U = 4;
    const int64_t cpr64 = (a.K + 255) / 256, total64 = a.N * cpr64;
    SDNQ_REQUIRE(total64 < (int64_t(1) << 30), SDNQ_EUNSUPPORTED, "weight too large");
    const int cpr = static_cast<int>(cpr64), total = static_cast<int>(total64);
    const int64_t want = (total64 + int64_t(kWarps) * U - 1) / (int64_t(kWarps) * U);
    const char* dq_grid = getenv("SDNQ_B200_DQ_GRID");             // tuning knob (read per call): CTAs per SM of grid
    const int64_t cap = int64_t(num_sms()) * (dq_grid != nullptr && atoi(dq_grid) > 0 ? atoi(dq_grid) : 16);
    const unsigned grid = static_cast<unsigned>(want < cap ? (want > 0 ? want : 1) : cap);
    const bool plain = a.up == nullptr && a.hadamard == 0 && a.gather == nullptr && a.out_scale == 1.f;
    cudaError_t e = cudaSuccess;
    const bool fast_int = plain && a.f.kind == SDNQ_W_INT && !a.codebook && ((a.group32 & 7) == 0 || a.group32 >= a.K32);
    // flat path: int4 / int8 in byte storage, groups inside rows, dense scale rows
    const int64_t total_oct = a.N * a.K / 8;
    const bool grouped = a.gpr32 > 1 && a.group_shift >= 3 && a.K32 % a.group32 == 0 && a.row_stride32 == a.gpr32;
    const bool rowwise = a.gpr32 == 1 && a.row_stride32 == 1 && a.group32 >= a.K32;
    const bool fp8_plain = plain && a.f.kind == SDNQ_W_FP8_E4M3FN && !a.codebook && a.zp == nullptr && ((a.group32 & 7) == 0 || a.group32 >= a.K32);
    const bool flat = (fast_int || fp8_plain) && (a.f.bits == 4 || a.f.bits == 8) && a.f.word_bytes == 1 && a.K32 % 8 == 0 && total_oct < (int64_t(1) << 31) &&
                      (grouped || rowwise) && (reinterpret_cast<uintptr_t>(a.weight) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    // rotated 8-bit weights without SVD: tensor-core un-rotate
    bool rot8 = false;
    if constexpr (sizeof(OutT) == 2) {
        rot8 = a.gather == nullptr && a.out_scale == 1.f && a.hadamard != 0 && a.up == nullptr && !a.codebook && a.f.bits == 8 && a.f.word_bytes == 1 && a.K32 % 8 == 0 &&
               ((a.group32 & 3) == 0 || a.group32 >= a.K32) && (reinterpret_cast<uintptr_t>(a.weight) & 3) == 0 &&
               (reinterpret_cast<uintptr_t>(out) & 7) == 0 && getenv("SDNQ_B200_HADAMARD_BUTTERFLY") == nullptr;
    }
    if (rot8) {
        if constexpr (sizeof(OutT) == 2)
            e = launch_pdl(dequant_rot8_kernel<OutT, U>, dim3(grid), dim3(kThreads), 0, st, a, reinterpret_cast<OutT*>(out), cpr, total);
    } else if (flat) {
        const bool twos = a.f.bits == 8 && !a.f.is_unsigned;
        const float bias = twos ? 8388736.0f : 8388608.0f - static_cast<float>(a.f.int_offset);
        const uint32_t flip = twos ? 0x80808080u : 0u;
        const uint32_t opr = static_cast<uint32_t>(a.K32 / 8);
        const uint32_t magic = static_cast<uint32_t>((uint64_t(1) << 32) / opr);          // floor(2^32 / opr): quotient estimate is exact or one low
        const int shift = grouped ? a.group_shift - 3 : 0;
        const int64_t blocks64 = (total_oct + int64_t(kThreads) * U - 1) / (int64_t(kThreads) * U);
        const char* gm = getenv("SDNQ_B200_FLAT_GRID");            // tuning knobs (read per call): CTAs per SM, streaming stores
        const char* se = getenv("SDNQ_B200_FLAT_STREAMING");
        // measured on B200 (tools/flat_tune.py, profiles/r01_dequant_flat_tuning.md): 32 CTAs per SM of grid (short grid-stride
        // loops) reach 5.9-6.2 TB/s where 8 gave 5.1-5.3; evict-first stores add 1-2 % for 4-bit sources and cost 1 % for 8-bit ones
        const int streaming = se != nullptr ? atoi(se) : (a.f.bits == 4 ? 1 : 0);
        const int64_t capf = int64_t(num_sms()) * (gm != nullptr && atoi(gm) > 0 ? atoi(gm) : 32);
        const unsigned gridf = static_cast<unsigned>(blocks64 < capf ? blocks64 : capf);
        OutT* o = reinterpret_cast<OutT*>(out);
        const uint32_t tot = static_cast<uint32_t>(total_oct);
#define SDNQ_FLAT(BITS_, ZP_, ROW_)                                                                                              \
    e = launch_pdl(dequant_flat_kernel<BITS_, ZP_, ROW_, OutT, U>, dim3(gridf), dim3(kThreads), 0, st, a.weight, a.scale, a.zp, o, tot, \
                   shift, opr, magic, flip, bias, streaming)
        const bool z = a.zp != nullptr;
        if (fp8_plain) {
            if (rowwise) e = launch_pdl(dequant_flat_kernel<8, false, true, OutT, U, true>, dim3(gridf), dim3(kThreads), 0, st, a.weight, a.scale, a.zp, o, tot,
                                        shift, opr, magic, flip, bias, streaming);
            else e = launch_pdl(dequant_flat_kernel<8, false, false, OutT, U, true>, dim3(gridf), dim3(kThreads), 0, st, a.weight, a.scale, a.zp, o, tot,
                                shift, opr, magic, flip, bias, streaming);
        } else if (a.f.bits == 4) {
            if (rowwise) { if (z) SDNQ_FLAT(4, true, true); else SDNQ_FLAT(4, false, true); }
            else { if (z) SDNQ_FLAT(4, true, false); else SDNQ_FLAT(4, false, false); }
        } else {
            if (rowwise) { if (z) SDNQ_FLAT(8, true, true); else SDNQ_FLAT(8, false, true); }
            else { if (z) SDNQ_FLAT(8, true, false); else SDNQ_FLAT(8, false, false); }
        }
#undef SDNQ_FLAT
    } else if (fast_int) {
        const bool pow2 = a.group_shift >= 0;
        if (a.zp != nullptr) {
            if (pow2) { SDNQ_DISPATCH_BITS(a.f.bits, (e = launch_pdl(dequant_int_kernel<BITS, true, true, OutT, U>, dim3(grid), dim3(kThreads), 0, st, a, reinterpret_cast<OutT*>(out), cpr, total))); }
            else { SDNQ_DISPATCH_BITS(a.f.bits, (e = launch_pdl(dequant_int_kernel<BITS, true, false, OutT, U>, dim3(grid), dim3(kThreads), 0, st, a, reinterpret_cast<OutT*>(out), cpr, total))); }
        } else {
            if (pow2) { SDNQ_DISPATCH_BITS(a.f.bits, (e = launch_pdl(dequant_int_kernel<BITS, false, true, OutT, U>, dim3(grid), dim3(kThreads), 0, st, a, reinterpret_cast<OutT*>(out), cpr, total))); }
            else { SDNQ_DISPATCH_BITS(a.f.bits, (e = launch_pdl(dequant_int_kernel<BITS, false, false, OutT, U>, dim3(grid), dim3(kThreads), 0, st, a, reinterpret_cast<OutT*>(out), cpr, total))); }
        }
    } else if (plain) {
        SDNQ_DISPATCH_BITS(a.f.bits, (e = launch_pdl(dequant_kernel<BITS, OutT, U, true>, dim3(grid), dim3(kThreads), 0, st, a, reinterpret_cast<OutT*>(out), cpr, total)));
    } else {
        SDNQ_DISPATCH_BITS(a.f.bits, (e = launch_pdl(dequant_kernel<BITS, OutT, U, false>, dim3(grid), dim3(kThreads), 0, st, a, reinterpret_cast<OutT*>(out), cpr, total)));
    }
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of dequant_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("dequant_kernel");
}

}  // namespace sdnq

using namespace sdnq;

extern "C" int sdnq_b200_unpack(const void* packed, const sdnq_weight_format* fmt, void* out, int out_dtype, int64_t numel,
                                void* stream) {
    WFormat f;
    int rc = make_wformat(fmt, &f);
    if (rc != SDNQ_OK) return rc;
    SDNQ_REQUIRE(packed && out, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(numel >= 0 && numel % 8 == 0, SDNQ_EINVAL, "numel (=%lld) must be a non-negative multiple of 8", (long long)numel);
    const bool is_float = f.kind != SDNQ_W_INT;
    if (is_float)
        SDNQ_REQUIRE(out_dtype == SDNQ_F32 || out_dtype == SDNQ_BF16 || out_dtype == SDNQ_F16 || out_dtype == SDNQ_F8E4M3, SDNQ_EINVAL,
                     "float formats unpack to f32/bf16/f16/f8e4m3");
    else
        SDNQ_REQUIRE(out_dtype == SDNQ_I8 || out_dtype == SDNQ_U8 || out_dtype == SDNQ_I32 || out_dtype == SDNQ_F32 ||
                         out_dtype == SDNQ_BF16 || out_dtype == SDNQ_F16, SDNQ_EINVAL, "bad out_dtype %d for an integer format", out_dtype);
    if (numel == 0) return SDNQ_OK;
    const int64_t octets = numel / 8;
    const unsigned blocks = static_cast<unsigned>((octets + kThreads - 1) / kThreads);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SDNQ_DISPATCH_BITS(f.bits, launch_plain(unpack_kernel<BITS>, dim3(blocks), dim3(kThreads), 0, st, reinterpret_cast<const uint8_t*>(packed), f, out, out_dtype, octets));
    launch_plain(operand_fence_kernel, dim3(1), dim3(32), 0, st);
    return check_launch("unpack_kernel");
}

extern "C" int sdnq_b200_dequant(const void* weight, const sdnq_weight_format* fmt, const float* scale, const float* zero_point,
                                 int codebook, int64_t N, int64_t K, int64_t group_size, const void* svd_up, int64_t up_stride_n,
                                 int64_t up_stride_r, const void* svd_down, int64_t down_stride_r, int64_t down_stride_k,
                                 int svd_rank, int svd_dtype, int hadamard_group, void* out, int out_dtype, void* stream) {
    DequantArgs a;
    int rc = fill_args(a, weight, fmt, scale, zero_point, codebook, N, K, group_size);
    if (rc != SDNQ_OK) return rc;
    SDNQ_REQUIRE(out != nullptr && (reinterpret_cast<uintptr_t>(out) & 15) == 0, SDNQ_EINVAL, "out must be a 16-byte aligned pointer");
    SDNQ_REQUIRE(hadamard_ok(hadamard_group), SDNQ_EUNSUPPORTED, "hadamard group %d: only powers of two in [4,256] are implemented", hadamard_group);
    if (hadamard_group) SDNQ_REQUIRE(K % hadamard_group == 0, SDNQ_EINVAL, "hadamard group %d does not divide K", hadamard_group);
    if (svd_up != nullptr) {
        SDNQ_REQUIRE(svd_down != nullptr && svd_rank > 0 && svd_rank <= 1024, SDNQ_EINVAL, "bad svd arguments (rank %d)", svd_rank);
        SDNQ_REQUIRE(svd_dtype == SDNQ_BF16 || svd_dtype == SDNQ_F16 || svd_dtype == SDNQ_F32, SDNQ_EINVAL, "bad svd dtype");
        a.up = svd_up; a.up_sn = up_stride_n; a.up_sr = up_stride_r;
        a.down = svd_down; a.down_sr = down_stride_r; a.down_sk = down_stride_k;
        a.rank = svd_rank; a.svd_dtype = svd_dtype;
    }
    a.hadamard = hadamard_group;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const char* svd_env = getenv("SDNQ_B200_SVD_TC");          // "0": keep the CUDA-core SVD update (A/B measurements), read per call
    if (svd_up != nullptr && hadamard_group == 0 && !codebook && !(svd_env != nullptr && svd_env[0] == '0')) {
        // rank-r correction on the tensor cores when the layout allows it (dequant_svd.cu); 1 = "not covered, use the generic kernel"
        rc = dequant_svd_tc(weight, a.f, scale, zero_point, N, K, a.group32, a.group_shift, a.gpr32, a.row_stride32, svd_up, up_stride_n, up_stride_r,
                            svd_down, down_stride_r, down_stride_k, svd_rank, svd_dtype, out, out_dtype, st);
        if (rc != 1) return rc;
    }
    switch (out_dtype) {
        case SDNQ_BF16: return launch_dequant<__nv_bfloat16>(a, out, st);
        case SDNQ_F16: return launch_dequant<__half>(a, out, st);
        case SDNQ_F32: return launch_dequant<float>(a, out, st);
        default: return set_error(SDNQ_EINVAL, "bad out_dtype %d", out_dtype);
    }
}

// ---- batched dequantisation (K3s over several layers' weights in one launch)
extern "C" size_t sdnq_b200_dequant_batch_table_bytes(int n_jobs) { return n_jobs > 0 ? size_t(n_jobs) * svd_batch_entry_bytes() : 0; }

extern "C" int sdnq_b200_dequant_batch_plan(const sdnq_dequant_job* jobs, int n_jobs, void* host_table, int32_t* info) {
    SDNQ_REQUIRE(jobs && host_table && info && n_jobs > 0, SDNQ_EINVAL, "NULL pointer or empty batch");
    // tile width: narrow tiles keep several CTAs per SM busy on the latency chain TMA -> MMA -> TMEM -> epilogue (SDNQ_B200_SVD_BATCH_TN forces it)
    int tn = 64;
    if (const char* e = getenv("SDNQ_B200_SVD_BATCH_TN")) {
        const int v = atoi(e);
        if (v == 64 || v == 128 || v == 256) tn = v;
    }
    int total = 0, max_rank = 0, traits = 0;
    uint8_t* entry = reinterpret_cast<uint8_t*>(host_table);
    for (int j = 0; j < n_jobs; ++j, entry += svd_batch_entry_bytes()) {
        const sdnq_dequant_job& q = jobs[j];
        DequantArgs a;
        int rc = fill_args(a, q.weight, &q.fmt, q.scale, q.zero_point, 0, q.N, q.K, q.group_size);
        if (rc != SDNQ_OK) return rc;
        SDNQ_REQUIRE(q.out != nullptr && (reinterpret_cast<uintptr_t>(q.out) & 15) == 0, SDNQ_EINVAL, "job %d: out must be a 16-byte aligned pointer", j);
        SDNQ_REQUIRE(q.svd_up != nullptr && q.svd_down != nullptr, SDNQ_EUNSUPPORTED, "job %d: batched dequantisation covers weights with SVD factors", j);
        int tiles = 0;
        rc = svd_batch_fill(entry, tn, total, q.weight, a.f, q.scale, q.zero_point, q.N, q.K, a.group32, a.group_shift, a.gpr32, a.row_stride32, q.svd_up,
                            q.up_stride_n, q.up_stride_r, q.svd_down, q.down_stride_r, q.down_stride_k, q.svd_rank, q.svd_dtype, q.out, q.out_dtype, &tiles, &traits);
        if (rc == 1) return set_error(SDNQ_EUNSUPPORTED, "job %d: outside what the batched tensor-core dequant kernel covers (int4 / uint4 codes, bf16 SVD "
                                      "factors of rank 16 / 32 / 64 stored K-major, bf16 output, K %% 32 == 0)", j);
        if (rc != SDNQ_OK) return rc;
        total += tiles;
        max_rank = q.svd_rank > max_rank ? q.svd_rank : max_rank;
    }
    info[0] = tn | (traits << 16); info[1] = total; info[2] = max_rank; info[3] = n_jobs;
    return SDNQ_OK;
}

extern "C" int sdnq_b200_dequant_batch_run(const void* device_table, const int32_t* info, void* stream) {
    SDNQ_REQUIRE(device_table && info && (reinterpret_cast<uintptr_t>(device_table) & 127) == 0, SDNQ_EINVAL, "device_table must be a 128-byte aligned device pointer");
    const int tn = info[0] & 0xFFFF, traits = info[0] >> 16;
    SDNQ_REQUIRE((tn == 64 || tn == 128 || tn == 256) && info[1] > 0 && info[2] > 0 && info[3] > 0, SDNQ_EINVAL, "bad plan info");
    return svd_batch_run(device_table, info[3], info[1], tn, info[2], traits, reinterpret_cast<cudaStream_t>(stream));
}

// ---- quantized embedding lookup: gather + dequantise the selected rows (layers/embedding/forward.py:14-68)
extern "C" int sdnq_b200_embedding(const void* weight, const sdnq_weight_format* fmt, const float* scale, const float* zero_point, int codebook,
                                   int64_t V, int64_t D, int64_t group_size, const void* svd_up, int64_t up_stride_n, int64_t up_stride_r,
                                   const void* svd_down, int64_t down_stride_r, int64_t down_stride_k, int svd_rank, int svd_dtype,
                                   int hadamard_group, const int64_t* indices, int64_t n_indices, float embed_scale, void* out, int out_dtype,
                                   void* stream) {
    DequantArgs a;
    int rc = fill_args(a, weight, fmt, scale, zero_point, codebook, V, D, group_size);
    if (rc != SDNQ_OK) return rc;
    SDNQ_REQUIRE(indices != nullptr && n_indices >= 0, SDNQ_EINVAL, "indices pointer is NULL");
    SDNQ_REQUIRE(out != nullptr && (reinterpret_cast<uintptr_t>(out) & 15) == 0, SDNQ_EINVAL, "out must be a 16-byte aligned pointer");
    SDNQ_REQUIRE(hadamard_ok(hadamard_group), SDNQ_EUNSUPPORTED, "hadamard group %d: only powers of two in [4,256] are implemented", hadamard_group);
    if (hadamard_group) SDNQ_REQUIRE(D % hadamard_group == 0, SDNQ_EINVAL, "hadamard group %d does not divide the embedding width", hadamard_group);
    SDNQ_REQUIRE(n_indices * D < (int64_t(1) << 31), SDNQ_EUNSUPPORTED, "too many looked-up elements (%lld x %lld)", (long long)n_indices, (long long)D);
    if (n_indices == 0) return SDNQ_OK;
    if (svd_up != nullptr) {
        SDNQ_REQUIRE(svd_down != nullptr && svd_rank > 0 && svd_rank <= 1024, SDNQ_EINVAL, "bad svd arguments (rank %d)", svd_rank);
        SDNQ_REQUIRE(svd_dtype == SDNQ_BF16 || svd_dtype == SDNQ_F16 || svd_dtype == SDNQ_F32, SDNQ_EINVAL, "bad svd dtype");
        a.up = svd_up; a.up_sn = up_stride_n; a.up_sr = up_stride_r;
        a.down = svd_down; a.down_sr = down_stride_r; a.down_sk = down_stride_k;
        a.rank = svd_rank; a.svd_dtype = svd_dtype;
    }
    a.hadamard = hadamard_group;
    a.gather = indices;
    a.src_rows = V;
    a.N = n_indices;                       // rows of the output
    a.out_scale = embed_scale;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (out_dtype) {
        case SDNQ_BF16: return launch_dequant<__nv_bfloat16>(a, out, st);
        case SDNQ_F16: return launch_dequant<__half>(a, out, st);
        case SDNQ_F32: return launch_dequant<float>(a, out, st);
        default: return set_error(SDNQ_EINVAL, "bad out_dtype %d", out_dtype);
    }
}

extern "C" int sdnq_b200_requant(const void* weight, const sdnq_weight_format* fmt, const float* scale, const float* zero_point,
                                 int codebook, int64_t N, int64_t K, int64_t group_size, int mm_dtype, void* wq, float* sw, float* zw,
                                 int32_t* colsum, void* stream) {
    DequantArgs a;
    int rc = fill_args(a, weight, fmt, scale, zero_point, codebook, N, K, group_size);
    if (rc != SDNQ_OK) return rc;
    SDNQ_REQUIRE(mm_dtype == SDNQ_I8 || mm_dtype == SDNQ_U8 || mm_dtype == SDNQ_F8E4M3, SDNQ_EINVAL, "bad matmul dtype %d", mm_dtype);
    SDNQ_REQUIRE(wq && sw && (mm_dtype != SDNQ_U8 || zw), SDNQ_EINVAL, "NULL output pointer");
    SDNQ_REQUIRE((reinterpret_cast<uintptr_t>(wq) & 7) == 0, SDNQ_EINVAL, "wq must be 8-byte aligned");
    SDNQ_REQUIRE(K <= int64_t(kRequantMaxOct) * 8 * kThreads, SDNQ_EUNSUPPORTED, "requant: K=%lld exceeds %d", (long long)K,
                 kRequantMaxOct * 8 * kThreads);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SDNQ_DISPATCH_BITS(a.f.bits, launch_plain(requant_kernel<BITS>, dim3(static_cast<unsigned>(N)), dim3(kThreads), 0, st,
                                              a, mm_dtype, reinterpret_cast<uint8_t*>(wq), sw, mm_dtype == SDNQ_U8 ? zw : nullptr, colsum));
    launch_plain(operand_fence_kernel, dim3(1), dim3(32), 0, st);
    return check_launch("requant_kernel");
}
