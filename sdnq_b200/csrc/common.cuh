// Shared host/device helpers for the sdnq_b200 kernels (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <utility>

#include "../../include/sdnq_b200.h"

namespace sdnq {

// ---------------------------------------------------------------- host: errors / launch accounting
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define SDNQ_REQUIRE(cond, code, ...)                                  \
    do {                                                               \
        if (!(cond)) return ::sdnq::set_error((code), __VA_ARGS__);    \
    } while (0)

#define SDNQ_CUDA_OK(expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return ::sdnq::set_error(SDNQ_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));   \
    } while (0)

#ifdef SDNQ_HOST_EMU
// tests/host_emu: the launch runs the kernel function on emulated CTAs (lock-stepped host threads), synchronously
inline int check_launch(const char*) {
    count_launch();
    return SDNQ_OK;
}
#else
inline int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(SDNQ_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    }
    count_launch();
    return SDNQ_OK;
}
#endif

int num_sms();
bool pdl_enabled();

// Launch with the programmatic-dependent-launch attribute: the kernel may start (prologue, weight prefetch) while its
// stream predecessor is still draining; every kernel of this library calls pdl_wait() before touching data a predecessor
// may have produced (and before exiting, which keeps the dependency chain transitive).
#ifdef SDNQ_HOST_EMU
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t, cudaStream_t, Args&&... args) {
    ::sdnq_emu::run_grid(grid, static_cast<int>(block.x), [&] { kernel(args...); });
    return cudaSuccess;
}
// an ordinary launch (kernel<<<grid, block, smem, stream>>>(args...)) without the programmatic-dependent-launch attribute
template <typename... KArgs, typename... Args>
void launch_plain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t, cudaStream_t, Args&&... args) {
    ::sdnq_emu::run_grid(grid, static_cast<int>(block.x), [&] { kernel(args...); });
}
#else
template <typename... KArgs, typename... Args>
void launch_plain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    kernel<<<grid, block, smem, st>>>(std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

// ---------------------------------------------------------------- device: programmatic dependent launch
#ifdef SDNQ_HOST_EMU
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_launch_dependents() {}
#else
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---------------------------------------------------------------- device: small numerics
__device__ __forceinline__ float bf16_to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> struct ElemTraits;
template <> struct ElemTraits<float> {
    static constexpr int kDtype = SDNQ_F32;
    __device__ static __forceinline__ float load(float v) { return v; }
    __device__ static __forceinline__ float round(float v) { return v; }
};
template <> struct ElemTraits<__nv_bfloat16> {
    static constexpr int kDtype = SDNQ_BF16;
    __device__ static __forceinline__ float load(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ static __forceinline__ float round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};
template <> struct ElemTraits<__half> {
    static constexpr int kDtype = SDNQ_F16;
    __device__ static __forceinline__ float load(__half v) { return __half2float(v); }
    __device__ static __forceinline__ float round(float v) { return __half2float(__float2half_rn(v)); }
};

// 8 consecutive elements of T held as floats <-> global memory (16 B for 2-byte types, 32 B for float)
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __half22float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = r;
}
template <>
__device__ __forceinline__ void store8<__half>(__half* p, const float (&v)[8]) {
    uint4 r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = r;
}

// 4 consecutive elements (half an octet)
template <typename T>
__device__ __forceinline__ void store4(T* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}
template <>
__device__ __forceinline__ void store4<__half>(__half* p, float a, float b, float c, float d) {
    __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

// float -> e4m3fn byte, round-to-nearest-even, input already clamped to +-448 (matches torch's cast)
__device__ __forceinline__ uint8_t f32_to_e4m3(float v) {
    return static_cast<uint8_t>(__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3));
}
__device__ __forceinline__ float e4m3_to_f32(uint8_t b) {
    __half_raw h = __nv_cvt_fp8_to_halfraw(static_cast<__nv_fp8_storage_t>(b), __NV_E4M3);
    return __half2float(*reinterpret_cast<__half*>(&h));
}
__device__ __forceinline__ float e5m2_to_f32(uint8_t b) {
    __half_raw h = __nv_cvt_fp8_to_halfraw(static_cast<__nv_fp8_storage_t>(b), __NV_E5M2);
    return __half2float(*reinterpret_cast<__half*>(&h));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- device: Hadamard over a warp
// A warp holds a 256-element chunk of one row: lane l owns elements [8l, 8l+8).  The transform is applied independently to
// each aligned sub-block of G elements (G = 4..256, power of two); the 1/sqrt(G) factor is applied by the caller.
//
//   G not a power of 4 : Sylvester kron^k([[1,1],[1,-1]])                                  (reference quant_utils.py:144-152)
//   G a power of 4     : kron^k(H4), H4 = [[1,1,1,-1],[1,1,-1,1],[1,-1,1,1],[-1,1,1,1]]    (reference quant_utils.py:155-165)
//
// Both run as radix-2 butterflies (3 in-register stages + one xor-shuffle stage per remaining bit).  For the H4 family we use
//   H4 = P . D . (H2 (x) H2) . D,   D = diag(1,1,1,-1),  P = swap of the two middle outputs,
// applied per base-4 digit: negate inputs whose digit is 3 (for every digit), run the Sylvester transform, negate outputs by
// the same rule, and deliver output p at position swap_bit_pairs(p).  The in-register half of that permutation (bits 0<->1)
// is done here by renaming registers; the cross-lane half is *not* moved through shuffles: callers store the two 4-element
// halves of the lane at hadamard_dest<G>(lane, half) instead (a permutation of 8 B / 4 B pieces inside the same 256-chunk).
template <int G> struct HadamardInfo {
    static_assert(G >= 4 && G <= 256 && (G & (G - 1)) == 0, "hadamard group must be a power of two in [4,256]");
    static constexpr int LOG = (G == 4) ? 2 : (G == 8) ? 3 : (G == 16) ? 4 : (G == 32) ? 5 : (G == 64) ? 6 : (G == 128) ? 7 : 8;
    static constexpr bool kPow4 = (LOG & 1) == 0;
};

// sign mask (0 or 0x80000000) of register j of this lane: product over the base-4 digits below LOG of (digit == 3 ? -1 : +1)
template <int G>
__device__ __forceinline__ uint32_t hadamard_sign_mask(int lane, int j) {
    constexpr int LOG = HadamardInfo<G>::LOG;
    bool neg = (j & 3) == 3;                                            // digit 0 = element bits 0-1
    if (LOG >= 4) neg ^= (j >= 4) && (lane & 1);                        // digit 1 = (lane bit 0, element bit 2)
    if (LOG >= 6) neg ^= ((lane >> 1) & 3) == 3;
    if (LOG >= 8) neg ^= ((lane >> 3) & 3) == 3;
    return neg ? 0x80000000u : 0u;
}

// In:  v = the lane's 8 elements.  Out: v = transformed elements times `factor` (1/sqrt(G) in the activation dtype), for
// power-of-4 groups with registers (1,2) and (5,6) exchanged and belonging at hadamard_dest<G>() (see above).
// The butterflies use the packed f32x2 pipe of sm_100 (FADD2 / FFMA2 / FMUL2: two floats per instruction).
template <int G>
__device__ __forceinline__ void hadamard_warp(float (&v)[8], float factor) {
    constexpr int LOG = HadamardInfo<G>::LOG;
    constexpr bool kPow4 = HadamardInfo<G>::kPow4;
    const int lane = threadIdx.x & 31;
    if constexpr (kPow4) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(__float_as_uint(v[j]) ^ hadamard_sign_mask<G>(lane, j));
    }
    float2 p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = make_float2(v[2 * i] + v[2 * i + 1], v[2 * i] - v[2 * i + 1]);      // element bit 0
    const float2 minus = make_float2(-1.f, -1.f);
    if constexpr (G >= 4) {                                                                               // element bit 1
#pragma unroll
        for (int i = 0; i < 4; i += 2) {
            const float2 a = p[i], b = p[i + 1];
            p[i] = __fadd2_rn(a, b);
            p[i + 1] = __ffma2_rn(b, minus, a);
        }
    }
    if constexpr (G >= 8) {                                                                               // element bit 2
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float2 a = p[i], b = p[i + 2];
            p[i] = __fadd2_rn(a, b);
            p[i + 2] = __ffma2_rn(b, minus, a);
        }
    }
#pragma unroll
    for (int b = 3; b < LOG; ++b) {                                                                       // element bits 3..: across lanes
        const int m = 1 << (b - 3);
        const float sg = (lane & m) ? -1.0f : 1.0f;       // upper half of the pair computes (other - mine)
        const float2 sgn = make_float2(sg, sg);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 o;
            o.x = __shfl_xor_sync(0xffffffffu, p[i].x, m);
            o.y = __shfl_xor_sync(0xffffffffu, p[i].y, m);
            p[i] = __ffma2_rn(sgn, p[i], o);
        }
    }
    // scale (and, for the H4 family, the output signs folded into the factor; then bits 0 <-> 1 of the output index)
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = kPow4 ? __uint_as_float(__float_as_uint(factor) ^ hadamard_sign_mask<G>(lane, j)) : factor;
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = __fmul2_rn(p[i], make_float2(f[2 * i], f[2 * i + 1]));
    if constexpr (kPow4) {
        v[0] = p[0].x; v[1] = p[1].x; v[2] = p[0].y; v[3] = p[1].y;
        v[4] = p[2].x; v[5] = p[3].x; v[6] = p[2].y; v[7] = p[3].y;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = p[i].x; v[2 * i + 1] = p[i].y; }
    }
}

// element offset inside the 256-chunk where registers [4*half, 4*half+4) of this lane belong after hadamard_warp<G>
template <int G>
__device__ __forceinline__ int hadamard_dest(int lane, int half) {
    constexpr int LOG = HadamardInfo<G>::LOG;
    int p = lane * 8 + half * 4;
    if constexpr (HadamardInfo<G>::kPow4 && LOG >= 4) {
        constexpr int mask = (1 << LOG) - 1;
        const int even = p & 0x54 & mask, odd = p & 0xA8 & mask;    // bit pairs (2,3), (4,5), (6,7) below LOG
        p = (p & ~(0xFC & mask)) | (even << 1) | (odd >> 1);
    }
    return p;
}

__device__ __forceinline__ void hadamard_warp_dyn(int G, float (&v)[8], float factor) {
    switch (G) {
        case 4: hadamard_warp<4>(v, factor); break;
        case 8: hadamard_warp<8>(v, factor); break;
        case 16: hadamard_warp<16>(v, factor); break;
        case 32: hadamard_warp<32>(v, factor); break;
        case 64: hadamard_warp<64>(v, factor); break;
        case 128: hadamard_warp<128>(v, factor); break;
        case 256: hadamard_warp<256>(v, factor); break;
        default: break;
    }
}
__device__ __forceinline__ int hadamard_dest_dyn(int G, int lane, int half) {
    switch (G) {
        case 16: return hadamard_dest<16>(lane, half);
        case 64: return hadamard_dest<64>(lane, half);
        case 256: return hadamard_dest<256>(lane, half);
        default: return lane * 8 + half * 4;            // no rotation, Sylvester groups and G = 4 keep their place
    }
}

// 1/sqrt(G) as the reference applies it: H.div_(n**0.5) in the activation dtype, so the factor is
// f32(1/sqrt(n)) rounded to T (exact power of two for G a power of 4).
template <typename T>
__device__ __forceinline__ float hadamard_factor(int G) {
    return ElemTraits<T>::round(__fdiv_rn(1.0f, sqrtf(static_cast<float>(G))));
}

}  // namespace sdnq
