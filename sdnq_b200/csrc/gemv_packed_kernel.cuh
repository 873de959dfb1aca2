// K5p: small-M Linear straight from the *stored* weight -- packed sub-byte integers, minifloats, fp8 / int8, row-wise or
// group-wise scales, optional zero points -- without materialising the dequantised weight ("W4A16 GEMV" and friends).
//
// Reference behaviour restated here: for fewer than 32 rows every forward dequantises the whole weight and calls F.linear
// (layers/linear/forward.py:24-26 for use_quantized_matmul=False; linear_int8.py:102-103, linear_uint8.py:107-108,
// linear_fp8.py:83-84 otherwise):
//     W[n,k] = cast_T( q[n,k] * s[n, k/g] )   or   cast_T( fma(q, s, zp) )          (dequantizer.py:15-84, f32 math, result dtype T)
//     y      = x @ W^T + bias
// i.e. 2 + bits/8 bytes of HBM traffic per weight element per call, for AdaLN / time-embedding Linears with M = batch.  This
// kernel reads the bits/8 bytes once: every lane unpacks its 16 codes with the decoders of unpack.cuh, applies scale / zero
// point in f32 in the reference's order, rounds to T -- the very values the reference's dequantised weight holds -- and feeds
// them to mma.sync.m16n8k16 (f32 accumulate) as the A operand; the activations are the B operand.  Same fragment mapping as
// K5 (gemv_w8a16.cu): weight rows are the MMA's M dimension (16 per CTA tile), K is split over the CTA's 8 warps and reduced
// through shared memory, activation rows are N (8 per block, up to 4 blocks).
//
// The body is a __device__ function template so that tests/host_emu can run it on a CPU (256 lock-stepped host threads with
// host models of the warp primitives) against the oracle; the __global__ wrapper lives in gemv_packed.cu.
#pragma once
#include "hadamard_tc.cuh"     // hadtc::Half16<T>: cvt pack + mma.sync m16n8k16 (with their host-emulation models)
#include "unpack.cuh"

namespace sdnq {
namespace gemvp {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct Args {
    const void* x;            // [M, K] activation dtype T, row stride ldx
    int64_t ldx;
    const uint8_t* w;         // stored weight: packed octets along the flattened [N, K] tensor (or plain 1-byte codes)
    const float* scale;       // [N, K / group]
    const float* zp;          // same shape or NULL
    const void* bias;         // NULL, [N] (bias_ld = 0) or [M, N] (bias_ld = row stride)
    int bias_dtype;
    int64_t bias_ld;
    void* out;                // [M, N] of T
    int M, N, K;
    int group;                // elements per scale along K (K for row-wise scales); a multiple of 8
    int groups_per_row;       // K / group
    WFormat f;
};

template <typename T>
__device__ __forceinline__ void store_out(void* p, int64_t i, float v);
template <>
__device__ __forceinline__ void store_out<__nv_bfloat16>(void* p, int64_t i, float v) { reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ void store_out<__half>(void* p, int64_t i, float v) { reinterpret_cast<__half*>(p)[i] = __float2half_rn(v); }

__device__ __forceinline__ float bias_at(const Args& a, int m, int n) {
    if (a.bias == nullptr) return 0.f;
    const int64_t i = int64_t(m) * a.bias_ld + n;
    if (a.bias_dtype == SDNQ_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.bias)[i]);
    if (a.bias_dtype == SDNQ_F16) return __half2float(reinterpret_cast<const __half*>(a.bias)[i]);
    return reinterpret_cast<const float*>(a.bias)[i];
}

// The 16 dequantised weights W[n, k .. k+16) of one lane as four A-fragment register pairs: p[2j] = (W[4j], W[4j+1]),
// p[2j+1] = (W[4j+2], W[4j+3]) rounded to T.  `live` = false gives zeros (rows past N, columns past K).
template <typename T, int BITS>
__device__ __forceinline__ void weights16(const Args& a, bool live, int n, int k, uint32_t (&p)[8]) {
    if (!live) {
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = 0u;
        return;
    }
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        const int ko = k + 8 * o;
        float q[8];
        uint32_t codes[8];
        octet_values<BITS>(a.w, (int64_t(n) * a.K + ko) >> 3, a.f, q, codes);
        const int64_t si = int64_t(n) * a.groups_per_row + (a.groups_per_row > 1 ? ko / a.group : 0);
        const float s = a.scale[si];
        float w[8];
        if (a.zp != nullptr) {
            const float z = a.zp[si];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = fmaf(q[i], s, z);          // addcmul(zp, q, scale)      dequantizer.py:15-48
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = q[i] * s;                  // q.to(f32) * scale          dequantizer.py:52-84
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) p[4 * o + i] = hadtc::Half16<T>::pack(w[2 * i], w[2 * i + 1]);
    }
}

// MB = number of 8-row activation blocks (M <= 8 * MB).  s_red: [kWarps - 1][MB * 4][32] floats of shared memory.
template <typename T, int BITS, int MB>
__device__ __forceinline__ void body(const Args& a, float* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const T* x = reinterpret_cast<const T*>(a.x);
    const int tiles = (a.N + 15) / 16;
    const int steps = (a.K + 63) / 64;                       // 64-column steps; in the last one lanes whose 16 columns lie past K idle
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int n_lo = tile * 16 + g, n_hi = n_lo + 8;
        const bool lo_ok = n_lo < a.N, hi_ok = n_hi < a.N;
        float acc[MB][4];
#pragma unroll
        for (int b = 0; b < MB; ++b) acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.f;
        for (int s = warp; s < steps; s += kWarps) {
            const int k = s * 64 + 16 * t;
            const bool live = k < a.K;                       // K % 16 == 0: a lane's 16 columns are all inside or all outside
            uint32_t wl[8], wh[8];
            weights16<T, BITS>(a, live && lo_ok, n_lo, k, wl);       // row g   : A regs a0 (k-slots 2t, 2t+1), a2 (2t+8, 2t+9)
            weights16<T, BITS>(a, live && hi_ok, n_hi, k, wh);       // row g+8 : A regs a1, a3
#pragma unroll
            for (int b = 0; b < MB; ++b) {
                uint4 x0 = make_uint4(0u, 0u, 0u, 0u), x1 = x0;        // the same 16 columns of activation row m = 8b + g
                if (live && 8 * b + g < a.M) {
                    const T* xr = x + int64_t(8 * b + g) * a.ldx + k;
                    x0 = *reinterpret_cast<const uint4*>(xr);
                    x1 = *reinterpret_cast<const uint4*>(xr + 8);
                }
                hadtc::Half16<T>::mma(acc[b], wl[0], wh[0], wl[1], wh[1], x0.x, x0.y);
                hadtc::Half16<T>::mma(acc[b], wl[2], wh[2], wl[3], wh[3], x0.z, x0.w);
                hadtc::Half16<T>::mma(acc[b], wl[4], wh[4], wl[5], wh[5], x1.x, x1.y);
                hadtc::Half16<T>::mma(acc[b], wl[6], wh[6], wl[7], wh[7], x1.z, x1.w);
            }
        }
        // ---- add the K-split partials up in warp 0
        if (warp > 0) {
#pragma unroll
            for (int b = 0; b < MB; ++b)
#pragma unroll
                for (int i = 0; i < 4; ++i) s_red[((warp - 1) * (MB * 4) + b * 4 + i) * 32 + lane] = acc[b][i];
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int w = 0; w < kWarps - 1; ++w)
#pragma unroll
                for (int b = 0; b < MB; ++b)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[b][i] += s_red[(w * (MB * 4) + b * 4 + i) * 32 + lane];
        }
        __syncthreads();                                     // s_red is reused by the next tile
        if (warp != 0) continue;
        // ---- epilogue: C fragment = (weight row g | g+8) x (activation rows 2t, 2t+1 of block b)
#pragma unroll
        for (int b = 0; b < MB; ++b) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int m = 8 * b + 2 * t + i;
                if (m >= a.M) continue;
                if (lo_ok) store_out<T>(a.out, int64_t(m) * a.N + n_lo, acc[b][i] + bias_at(a, m, n_lo));
                if (hi_ok) store_out<T>(a.out, int64_t(m) * a.N + n_hi, acc[b][2 + i] + bias_at(a, m, n_hi));
            }
        }
    }
}

}  // namespace gemvp
}  // namespace sdnq
