// K5: small-M Linear on 8-bit weights without materialising the dequantised weight ("W8A16 GEMV").
//
// Reference behaviour restated here: every quantized-matmul forward falls back to dequantise + F.linear when the input has
// fewer than 32 rows (layers/linear/linear_int8.py:102-103, linear_uint8.py:107-108, linear_fp8.py:83-84):
//     y = x @ dequant(W)^T + bias,        dequant(W)[n,k] = q[n,k] * s[n] (+ zp[n]),  row-wise scales
// i.e. AdaLN / time-embedding Linears (M = batch).  The reference (and our K3 + library GEMM path) writes the full bf16 weight
// to HBM and reads it back for every call: 3 bytes of traffic per weight instead of 1, plus an M-row GEMM the library runs at a
// few percent of peak.  Here the 1-byte codes are read once and multiplied on the fly:
//     y[m,n] = s[n] * sum_k x[m,k] q[n,k]  (+ zp[n] * sum_k x[m,k])  + bias[n]
// Products and sums are exact-or-f32 (the codes are exact in bf16 / f16, accumulation in f32), the scale is applied once per
// output -- slightly *more* accurate than the reference, which rounds every q*s to the activation dtype first; parity is held
// to the same tolerance as the other paths that contain a 16-bit GEMM (tests/test_layers_gpu.py).
// Rotated layers (use_hadamard): y = x @ (W_rot H)^T = (x H) @ W_rot^T, so the caller passes x already rotated (K2's x_rot).
//
// Mapping: the contraction runs on the tensor cores as mma.sync.m16n8k16 with the *weight rows* as the M dimension (16 rows per
// CTA tile, K split over the CTA's 8 warps and reduced through shared memory) and the activation rows as N (8 per block, up to 4 blocks = 32 rows).  Lane (g, t) streams 16 consecutive codes of
// rows g and g+8 per 64-column step (two 16-byte loads, 4 steps in flight), converts them to the activation dtype in registers
// (exact) and feeds 4 MMAs; the k-slots of the fragments are permuted so that these 16 codes are exactly what the lane needs, and
// the matching 16 activations of row m = g come from one 32-byte read of x (L1-resident: x is M*K*2 bytes).  HBM-bound on the
// codes: algorithmic bytes N*K (+ 2*M*K + 2*M*N).
#include "common.cuh"

namespace sdnq {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct GemvArgs {
    const void* x;          // [M, K] activation dtype, row stride ldx
    int64_t ldx;
    const uint8_t* wq;      // [N, K] 1-byte codes
    const float* sw;        // [N]
    const float* zp;        // [N] or NULL
    const void* bias;       // [N] or NULL
    int bias_dtype;
    void* out;              // [M, N]
    int M, N, K;
};

template <typename T> struct Act;
template <> struct Act<__nv_bfloat16> {
    __device__ static __forceinline__ uint32_t pack(float lo, float hi) {
#ifdef SDNQ_HOST_EMU
        return ::sdnq_emu::pack16x2(false, lo, hi);
#else
        uint32_t r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
#endif
    }
    __device__ static __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef SDNQ_HOST_EMU
        ::sdnq_emu::mma_m16n8k16(false, d, a[0], a[1], a[2], a[3], b0, b1);
#else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
#endif
    }
    __device__ static __forceinline__ float lo(uint32_t w) { return __uint_as_float(w << 16); }
    __device__ static __forceinline__ float hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
    __device__ static __forceinline__ void store(void* p, int64_t i, float v) { reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v); }
};
template <> struct Act<__half> {
    __device__ static __forceinline__ uint32_t pack(float lo, float hi) {
#ifdef SDNQ_HOST_EMU
        return ::sdnq_emu::pack16x2(true, lo, hi);
#else
        uint32_t r;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
#endif
    }
    __device__ static __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef SDNQ_HOST_EMU
        ::sdnq_emu::mma_m16n8k16(true, d, a[0], a[1], a[2], a[3], b0, b1);
#else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
#endif
    }
    __device__ static __forceinline__ float lo(uint32_t w) { return __half2float(__ushort_as_half(static_cast<unsigned short>(w & 0xFFFFu))); }
    __device__ static __forceinline__ float hi(uint32_t w) { return __half2float(__ushort_as_half(static_cast<unsigned short>(w >> 16))); }
    __device__ static __forceinline__ void store(void* p, int64_t i, float v) { reinterpret_cast<__half*>(p)[i] = __float2half_rn(v); }
};

// four consecutive codes (one 32-bit word) -> two packed pairs of T, exact
template <typename T, int kFp8>      // 0: int8 codes, 1: float8_e4m3fn, 2: float8_e5m2
__device__ __forceinline__ void codes4(uint32_t w, uint32_t& p01, uint32_t& p23) {
    if constexpr (kFp8 != 0) {
        // e4m3 pair -> f16 pair (exact, one instruction); f16 needs nothing more, bf16 goes through f32 (exact: 4 significant bits).
        // e5m2 is the upper byte of an f16: a byte permute.
        uint32_t h01, h23;
        if constexpr (kFp8 == 2) {
            h01 = __byte_perm(w, 0u, 0x1404);
            h23 = __byte_perm(w, 0u, 0x3424);
        } else {
#ifdef SDNQ_HOST_EMU
        h01 = ::sdnq_emu::e4m3x2_to_f16x2(static_cast<unsigned short>(w & 0xFFFFu));
        h23 = ::sdnq_emu::e4m3x2_to_f16x2(static_cast<unsigned short>(w >> 16));
#else
        asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h01) : "h"(static_cast<unsigned short>(w & 0xFFFFu)));
        asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h23) : "h"(static_cast<unsigned short>(w >> 16)));
#endif
        }
        if constexpr (ElemTraits<T>::kDtype == SDNQ_F16) {
            p01 = h01;
            p23 = h23;
        } else {
            const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
            const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
            p01 = Act<T>::pack(f01.x, f01.y);
            p23 = Act<T>::pack(f23.x, f23.y);
        }
    } else {
        // int8 -> f32 (byte-select I2F) -> pack (exact: |q| <= 128 has 8 significant bits)
        const float c0 = static_cast<float>(static_cast<int8_t>(w & 0xFFu)), c1 = static_cast<float>(static_cast<int8_t>((w >> 8) & 0xFFu));
        const float c2 = static_cast<float>(static_cast<int8_t>((w >> 16) & 0xFFu)), c3 = static_cast<float>(static_cast<int8_t>(w >> 24));
        p01 = Act<T>::pack(c0, c1);
        p23 = Act<T>::pack(c2, c3);
    }
}

// MB = number of 8-row activation blocks (M <= 8 * MB)
template <typename T, int kFp8, int MB>
__global__ void __launch_bounds__(kThreads) gemv_w8a16_kernel(const GemvArgs a) {
    __shared__ float s_xsum[32];
    __shared__ float s_red[kWarps - 1][MB * 4][32];          // partial accumulators of warps 1..7 (the K split of a tile)
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const T* x = reinterpret_cast<const T*>(a.x);
    if (a.zp != nullptr) {                                   // sum_k x[m,k] for the zero-point term (x is tiny and cache-resident)
        for (int m = warp; m < a.M; m += kWarps) {
            float s = 0.f;
            for (int k = lane; k < a.K; k += 32) s += ElemTraits<T>::load(x[int64_t(m) * a.ldx + k]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) s_xsum[m] = s;
        }
        __syncthreads();
    }
    const int tiles = (a.N + 15) / 16;
    const int steps = a.K / 64;                              // whole 64-column steps; the tail (K % 64, a multiple of 16) is handled below
    const int tail16 = (a.K - steps * 64) / 16;
    // One CTA per 16-row tile; its 8 warps split K (warp w takes the 64-column steps w, w+8, ...) and warp 0 adds the partial
    // accumulators up: every SM holds several CTAs, so >= 100 KB of codes are in flight per SM even when N is small.
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int n_lo = tile * 16 + g, n_hi = n_lo + 8;
        const bool lo_ok = n_lo < a.N, hi_ok = n_hi < a.N;
        const uint8_t* w_lo = a.wq + int64_t(lo_ok ? n_lo : 0) * a.K + 16 * t;
        const uint8_t* w_hi = a.wq + int64_t(hi_ok ? n_hi : 0) * a.K + 16 * t;
        float acc[MB][4];
#pragma unroll
        for (int b = 0; b < MB; ++b) acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.f;
        const T* xrow[MB];
        bool x_ok[MB];
#pragma unroll
        for (int b = 0; b < MB; ++b) {
            x_ok[b] = 8 * b + g < a.M;
            xrow[b] = x + int64_t(x_ok[b] ? 8 * b + g : 0) * a.ldx + 16 * t;
        }
        auto step = [&](const uint4& cl, const uint4& ch, int kbase) {
            // A fragments of the 4 MMAs of this 64-column step: MMA j takes codes [4j, 4j+4) of the lane's 16
            uint32_t af[4][4];
            const uint32_t wl[4] = {cl.x, cl.y, cl.z, cl.w}, wh[4] = {ch.x, ch.y, ch.z, ch.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                codes4<T, kFp8>(wl[j], af[j][0], af[j][2]);          // row g   : k-slots (2t, 2t+1), (2t+8, 2t+9)
                codes4<T, kFp8>(wh[j], af[j][1], af[j][3]);          // row g+8
            }
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                uint4 x0 = make_uint4(0u, 0u, 0u, 0u), x1 = x0;        // 16 activations of row m = 8b + g, same columns as the codes
                if (x_ok[b]) {
                    x0 = *reinterpret_cast<const uint4*>(xrow[b] + kbase);
                    x1 = *reinterpret_cast<const uint4*>(xrow[b] + kbase + 8);
                }
                Act<T>::mma(acc[b], af[0], x0.x, x0.y);
                Act<T>::mma(acc[b], af[1], x0.z, x0.w);
                Act<T>::mma(acc[b], af[2], x1.x, x1.y);
                Act<T>::mma(acc[b], af[3], x1.z, x1.w);
            }
        };
        constexpr int U = 4;                                 // steps in flight per warp
        int s = warp;
        for (; s + (U - 1) * kWarps < steps; s += U * kWarps) {
            uint4 cl[U], ch[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                cl[u] = *reinterpret_cast<const uint4*>(w_lo + (s + u * kWarps) * 64);
                ch[u] = *reinterpret_cast<const uint4*>(w_hi + (s + u * kWarps) * 64);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) step(cl[u], ch[u], (s + u * kWarps) * 64);
        }
        for (; s < steps; s += kWarps) step(*reinterpret_cast<const uint4*>(w_lo + s * 64), *reinterpret_cast<const uint4*>(w_hi + s * 64), s * 64);
        if (tail16 > 0 && warp == kWarps - 1) {                                    // K % 64 in {16, 32, 48}: lanes with t < tail16 hold real columns
            uint4 cl = make_uint4(0u, 0u, 0u, 0u), ch = cl;
            const bool live = t < tail16;
            if (live) {
                cl = *reinterpret_cast<const uint4*>(w_lo + steps * 64);
                ch = *reinterpret_cast<const uint4*>(w_hi + steps * 64);
            }
            uint32_t af[4][4];
            const uint32_t wl[4] = {cl.x, cl.y, cl.z, cl.w}, wh[4] = {ch.x, ch.y, ch.z, ch.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                codes4<T, kFp8>(wl[j], af[j][0], af[j][2]);
                codes4<T, kFp8>(wh[j], af[j][1], af[j][3]);
            }
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                uint4 x0 = make_uint4(0u, 0u, 0u, 0u), x1 = x0;
                if (x_ok[b] && live) {
                    x0 = *reinterpret_cast<const uint4*>(xrow[b] + steps * 64);
                    x1 = *reinterpret_cast<const uint4*>(xrow[b] + steps * 64 + 8);
                }
                Act<T>::mma(acc[b], af[0], x0.x, x0.y);
                Act<T>::mma(acc[b], af[1], x0.z, x0.w);
                Act<T>::mma(acc[b], af[2], x1.x, x1.y);
                Act<T>::mma(acc[b], af[3], x1.z, x1.w);
            }
        }
        // ---- add the K-split partials up in warp 0
        if (warp > 0) {
#pragma unroll
            for (int b = 0; b < MB; ++b)
#pragma unroll
                for (int i = 0; i < 4; ++i) s_red[warp - 1][b * 4 + i][lane] = acc[b][i];
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int w = 0; w < kWarps - 1; ++w)
#pragma unroll
                for (int b = 0; b < MB; ++b)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[b][i] += s_red[w][b * 4 + i][lane];
        }
        __syncthreads();                                     // s_red is reused by the next tile
        if (warp != 0) continue;
        // ---- epilogue: C fragment = (weight row g | g+8) x (activation rows 2t, 2t+1 of block b)
        const float s_lo = lo_ok ? a.sw[n_lo] : 0.f, s_hi = hi_ok ? a.sw[n_hi] : 0.f;
        const float z_lo = (a.zp != nullptr && lo_ok) ? a.zp[n_lo] : 0.f, z_hi = (a.zp != nullptr && hi_ok) ? a.zp[n_hi] : 0.f;
        auto bias_at = [&](int n) -> float {
            if (a.bias == nullptr) return 0.f;
            if (a.bias_dtype == SDNQ_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.bias)[n]);
            if (a.bias_dtype == SDNQ_F16) return __half2float(reinterpret_cast<const __half*>(a.bias)[n]);
            return reinterpret_cast<const float*>(a.bias)[n];
        };
        const float b_lo = lo_ok ? bias_at(n_lo) : 0.f, b_hi = hi_ok ? bias_at(n_hi) : 0.f;
#pragma unroll
        for (int b = 0; b < MB; ++b) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int m = 8 * b + 2 * t + i;
                if (m >= a.M) continue;
                const float xs = a.zp != nullptr ? s_xsum[m] : 0.f;
                if (lo_ok) Act<T>::store(a.out, int64_t(m) * a.N + n_lo, fmaf(acc[b][i], s_lo, fmaf(z_lo, xs, b_lo)));
                if (hi_ok) Act<T>::store(a.out, int64_t(m) * a.N + n_hi, fmaf(acc[b][2 + i], s_hi, fmaf(z_hi, xs, b_hi)));
            }
        }
    }
}

template <typename T, int kFp8>
int launch_mb(const GemvArgs& a, cudaStream_t st) {
    const int tiles = (a.N + 15) / 16;
    const int cap = num_sms() * 8;
    const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
    cudaError_t e;
    const int mb = (a.M + 7) / 8;
    if (mb <= 1) e = launch_pdl(gemv_w8a16_kernel<T, kFp8, 1>, dim3(grid), dim3(kThreads), 0, st, a);
    else if (mb == 2) e = launch_pdl(gemv_w8a16_kernel<T, kFp8, 2>, dim3(grid), dim3(kThreads), 0, st, a);
    else e = launch_pdl(gemv_w8a16_kernel<T, kFp8, 4>, dim3(grid), dim3(kThreads), 0, st, a);
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of gemv_w8a16_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("gemv_w8a16_kernel");
}

}  // namespace
}  // namespace sdnq

using namespace sdnq;

extern "C" int sdnq_b200_linear_small_m(const void* x, int x_dtype, int64_t ldx, const void* wq, int w_dtype, const float* sw, const float* zp,
                                        const void* bias, int bias_dtype, void* out, int64_t M, int64_t N, int64_t K, void* stream) {
    SDNQ_REQUIRE(x && wq && sw && out, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(M >= 0 && M <= 32 && N > 0 && K > 0, SDNQ_EINVAL, "small-M Linear: 0 <= M <= 32 (got M=%lld N=%lld K=%lld)", (long long)M, (long long)N, (long long)K);
    SDNQ_REQUIRE(K % 16 == 0 && ldx % 8 == 0 && ldx >= K, SDNQ_EUNSUPPORTED, "small-M Linear: K %% 16 == 0 and ldx %% 8 == 0 (K=%lld ldx=%lld)", (long long)K, (long long)ldx);
    SDNQ_REQUIRE(w_dtype == SDNQ_I8 || w_dtype == SDNQ_F8E4M3 || w_dtype == SDNQ_F8E5M2, SDNQ_EINVAL,
                 "weight codes must be int8, float8_e4m3fn or float8_e5m2 (got %d)", w_dtype);
    SDNQ_REQUIRE(x_dtype == SDNQ_BF16 || x_dtype == SDNQ_F16, SDNQ_EUNSUPPORTED, "small-M Linear: bf16 / f16 activations (got %d)", x_dtype);
    SDNQ_REQUIRE(bias == nullptr || bias_dtype == SDNQ_BF16 || bias_dtype == SDNQ_F16 || bias_dtype == SDNQ_F32, SDNQ_EINVAL, "bad bias dtype %d", bias_dtype);
    SDNQ_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(wq) & 15) == 0, SDNQ_EINVAL, "x and wq must be 16-byte aligned");
    SDNQ_REQUIRE(N * K < (int64_t(1) << 40) && N < (int64_t(1) << 31) && K < (int64_t(1) << 31), SDNQ_EUNSUPPORTED, "weight too large");
    if (M == 0) return SDNQ_OK;
    GemvArgs a{x, ldx, reinterpret_cast<const uint8_t*>(wq), sw, zp, bias, bias_dtype, out, int(M), int(N), int(K)};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (x_dtype == SDNQ_BF16)
        return w_dtype == SDNQ_F8E4M3 ? launch_mb<__nv_bfloat16, 1>(a, st) : w_dtype == SDNQ_F8E5M2 ? launch_mb<__nv_bfloat16, 2>(a, st) : launch_mb<__nv_bfloat16, 0>(a, st);
    return w_dtype == SDNQ_F8E4M3 ? launch_mb<__half, 1>(a, st) : w_dtype == SDNQ_F8E5M2 ? launch_mb<__half, 2>(a, st) : launch_mb<__half, 0>(a, st);
}
