// K8: load-time weight quantisation -- scale, round, clamp and pack in one pass over the weight.
//
// Reference behaviour restated here (the middle of sdnq_quantize_layer_weight, quantizer.py:236-253, for integer formats):
//   quantize_weight         quant_utils.py:27-56     scale = amax / qmax  |  (max - min) / (qmax - qmin), zero_point = min;
//                                                   q = round_half_even((w [- zero_point]) / scale), clamp(qmin, qmax)
//   get_scale_*             quant_utils.py:9-24      true f32 divisions (what the reference computes on the CPU; the fixtures in
//                                                   tests/golden were produced there)
//   pack_int                packed_int/__init__.py:76-80, pack.py:201-321   signed codes stored offset-binary, 8 codes per `bits` bytes
// The reference runs these as ~10 eager tensor ops per layer over an f32 copy of the weight; here a thread owns 8 consecutive
// weights (one octet of the packed layout), the scale group is reduced with warp shuffles (groups of 8..256 weights) or by the
// CTA (row-wise / wider groups, second read from L2), and the packed bytes are written directly: HBM traffic = the weight once
// in its own dtype + bits/8 bytes per weight out.
#include <cmath>

#include "act_quant.cuh"     // actq::RowDivider: correctly rounded x / s with the reciprocal hoisted out of the element loop
#include "unpack.cuh"

namespace sdnq {
namespace {

constexpr int kThreads = 256;

struct WQArgs {
    const void* w;          // [N*K] weights, row-major, dtype w_dtype (f32 / bf16 / f16)
    int w_dtype;
    int64_t octets;         // N*K / 8
    int oct_per_group;      // scale group size / 8
    int bits, is_unsigned, packed;      // packed = 0: one byte per code (int8 / uint8 storage), no offset
    int kind, exponent, mantissa;       // sdnq_wkind; float kinds: qmin / qmax are the format's smallest / largest value
    float qmin, qmax;
    int scale_dtype;        // SDNQ_F32: keep; SDNQ_BF16 / SDNQ_F16: round scale (and zero point) to that type first (dequantize_fp32=False)
    uint8_t* out;
    float* scale;           // [groups]
    float* zp;              // [groups] (unsigned formats)
};

// 8 unsigned codes -> the BITS storage bytes of their octet: the inverse of decode_octet (unpack.cuh)
template <int BITS>
__device__ __forceinline__ void encode_octet(const uint32_t (&v)[8], uint8_t (&b)[BITS]) {
    if constexpr (BITS == 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) b[i] = static_cast<uint8_t>(v[i]);
    } else if constexpr (BITS == 7) {
#pragma unroll
        for (int i = 0; i < 7; ++i) b[i] = static_cast<uint8_t>(v[i] | ((v[7] << (i + 1)) & 0x80u));
    } else if constexpr (BITS == 6) {
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int i = 0; i < 3; ++i) b[3 * g + i] = static_cast<uint8_t>(v[4 * g + i] | ((v[4 * g + 3] << (2 * (i + 1))) & 0xC0u));
    } else if constexpr (BITS == 5) {
        b[0] = static_cast<uint8_t>(v[0] | ((v[5] & 7u) << 5));
        b[1] = static_cast<uint8_t>(v[1] | ((v[6] & 7u) << 5));
        b[2] = static_cast<uint8_t>(v[2] | ((v[7] & 7u) << 5));
        b[3] = static_cast<uint8_t>(v[3] | (((v[5] >> 3) & 3u) << 5) | (((v[7] >> 4) & 1u) << 7));
        b[4] = static_cast<uint8_t>(v[4] | (((v[6] >> 3) & 3u) << 5) | (((v[7] >> 3) & 1u) << 7));
    } else if constexpr (BITS == 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = static_cast<uint8_t>(v[2 * i] | (v[2 * i + 1] << 4));
    } else if constexpr (BITS == 3) {
        b[0] = static_cast<uint8_t>(v[0] | (v[3] << 3) | ((v[6] & 3u) << 6));
        b[1] = static_cast<uint8_t>(v[1] | (v[4] << 3) | ((v[7] & 3u) << 6));
        b[2] = static_cast<uint8_t>(v[2] | (v[5] << 3) | (((v[6] >> 2) & 1u) << 6) | (((v[7] >> 2) & 1u) << 7));
    } else if constexpr (BITS == 2) {
        b[0] = static_cast<uint8_t>(v[0] | (v[1] << 2) | (v[2] << 4) | (v[3] << 6));
        b[1] = static_cast<uint8_t>(v[4] | (v[5] << 2) | (v[6] << 4) | (v[7] << 6));
    } else {
        uint32_t x = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) x |= v[j] << j;
        b[0] = static_cast<uint8_t>(x);
    }
}

__device__ __forceinline__ void load_octet_f32(const void* w, int w_dtype, int64_t oct, float (&v)[8]) {
    if (w_dtype == SDNQ_F32) load8<float>(reinterpret_cast<const float*>(w) + oct * 8, v);
    else if (w_dtype == SDNQ_BF16) load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(w) + oct * 8, v);
    else load8<__half>(reinterpret_cast<const __half*>(w) + oct * 8, v);
}

__device__ __forceinline__ float round_scale(float s, int scale_dtype) {
    if (scale_dtype == SDNQ_BF16) return __bfloat162float(__float2bfloat16_rn(s));
    if (scale_dtype == SDNQ_F16) return __half2float(__float2half_rn(s));
    return s;
}

// statistics of a group -> (scale, zero point) as the reference computes them
__device__ __forceinline__ void group_scale(const WQArgs& a, float amax, float vmin, float vmax, float& scale, float& zero) {
    if (a.is_unsigned) {
        scale = __fdiv_rn(__fsub_rn(vmax, vmin), __fsub_rn(a.qmax, a.qmin));
        zero = vmin;                                            // qmin == 0 for every unsigned format: zero_point = min
        if (a.qmin != 0.f) zero = __fsub_rn(vmin, __fmul_rn(scale, a.qmin));
        zero = round_scale(zero, a.scale_dtype);
    } else {
        scale = __fdiv_rn(amax, a.qmax);
        zero = 0.f;
    }
    scale = round_scale(scale, a.scale_dtype);
}

// pack_float (packed_float.py:26-82) on one clamped value: the reference's own bit arithmetic, which is not plain round-to-nearest --
// a normal is rounded up when the top four dropped mantissa bits exceed one half (strictly), a value below the smallest normal becomes
// round_half_even(|x| * 2^M / min_normal) mantissa steps, then sign / exponent / mantissa are squeezed into `bits` bits.
__device__ __forceinline__ uint32_t encode_minifloat(float x, int E, int M, int bits, int is_unsigned) {
    const int drop = 23 - M;
    int xi = static_cast<int>(__float_as_uint(x));
    const int top4 = (-(1 << (drop - 4))) & ~(-(1 << drop));
    if ((xi & top4) > (1 << (drop - 1))) xi += (1 << drop);
    if (E < 8) {
        const float min_normal = __uint_as_float(static_cast<uint32_t>(127 + 2 - (1 << (E - 1))) << 23);          // 2^(2 - 2^(E-1))
        const float ax = fabsf(__uint_as_float(static_cast<uint32_t>(xi)));
        if (ax < min_normal) {
            const float steps = rintf(__fmul_rn(ax, __fdiv_rn(static_cast<float>(1 << M), min_normal)));
            xi = (xi & static_cast<int>(0x80000000u)) | (static_cast<int>(steps) << drop);
        }
    }
    xi >>= drop;                                                // arithmetic, as torch's int32 shift
    const int sign_mask = is_unsigned ? (1 << (bits - 1)) : (1 << (bits - 1)) + (1 << (bits - 2));
    return static_cast<uint32_t>((((xi >> (8 - E)) & sign_mask) | (xi & ~sign_mask)) & ((1 << bits) - 1));
}

template <int BITS>
__device__ __forceinline__ void quantise_store(const WQArgs& a, int64_t oct, const float (&v)[8], float scale, float zero) {
    uint32_t codes[8];
    const int offset = (a.packed && !a.is_unsigned) ? static_cast<int>(a.qmin) : 0;      // packed signed codes are offset-binary
    const actq::RowDivider divider(scale);
    const bool safe = divider.safe();
    if (a.kind == SDNQ_W_INT) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float num = a.is_unsigned ? __fsub_rn(v[i], zero) : v[i];
            float q = safe ? divider.div<true>(num) : divider.div<false>(num);
            const bool nan = !(q == q);                         // 0 / 0 of an all-zero group: NaN survives round_ / clamp_, the int cast makes it 0
            q = fminf(fmaxf(rintf(q), a.qmin), a.qmax);
            const int c = nan ? 0 : static_cast<int>(q);
            codes[i] = static_cast<uint32_t>(c - offset) & 0xFFu;
        }
    } else {
        // float formats (quant_utils.py:47-55): nan_to_num_ (NaN -> 0, +-inf -> +-FLT_MAX), clamp to the format's range, then the cast
        // to torch.float8_e4m3fn / float8_e5m2 (round to nearest even) or pack_float's encoder for the eXmY minifloats
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float num = a.is_unsigned ? __fsub_rn(v[i], zero) : v[i];
            float q = safe ? divider.div<true>(num) : divider.div<false>(num);
            if (!(q == q)) q = 0.f;
            else if (q == 0.f) q = copysignf(0.f, num);         // the divider's residual step loses the sign of a zero quotient; x / s keeps it (-0 encodes as 0x80)
            q = fminf(fmaxf(q, a.qmin), a.qmax);                // (+-inf land on the range ends, as FLT_MAX would)
            if (a.kind == SDNQ_W_FP8_E4M3FN) codes[i] = __nv_cvt_float_to_fp8(q, __NV_SATFINITE, __NV_E4M3);
            else if (a.kind == SDNQ_W_FP8_E5M2) codes[i] = __nv_cvt_float_to_fp8(q, __NV_SATFINITE, __NV_E5M2);
            else codes[i] = encode_minifloat(q, a.exponent, a.mantissa, a.bits, a.is_unsigned);
        }
    }
    uint8_t bytes[BITS];
    encode_octet<BITS>(codes, bytes);
    uint8_t* dst = a.out + oct * BITS;
    if constexpr (BITS == 8) {
        *reinterpret_cast<uint2*>(dst) = make_uint2(uint32_t(bytes[0]) | (uint32_t(bytes[1]) << 8) | (uint32_t(bytes[2]) << 16) | (uint32_t(bytes[3]) << 24),
                                                   uint32_t(bytes[4]) | (uint32_t(bytes[5]) << 8) | (uint32_t(bytes[6]) << 16) | (uint32_t(bytes[7]) << 24));
    } else if constexpr (BITS == 4) {
        *reinterpret_cast<uint32_t*>(dst) = uint32_t(bytes[0]) | (uint32_t(bytes[1]) << 8) | (uint32_t(bytes[2]) << 16) | (uint32_t(bytes[3]) << 24);
    } else if constexpr (BITS == 2) {
        *reinterpret_cast<uint16_t*>(dst) = static_cast<uint16_t>(uint32_t(bytes[0]) | (uint32_t(bytes[1]) << 8));
    } else {
#pragma unroll
        for (int i = 0; i < BITS; ++i) dst[i] = bytes[i];
    }
}

// Groups of 8 .. 256 weights (a power of two): the group's octets sit in 1 .. 32 neighbouring lanes of one warp.
template <int BITS>
__global__ void __launch_bounds__(kThreads) weight_quant_group_kernel(const WQArgs a) {
    const int T = a.oct_per_group;                              // lanes per group (power of two <= 32)
    const int64_t rounded = (a.octets + 31) / 32 * 32;          // whole warps: the shuffles below need every lane
    for (int64_t oct = int64_t(blockIdx.x) * kThreads + threadIdx.x; oct < rounded; oct += int64_t(gridDim.x) * kThreads) {
        const bool ok = oct < a.octets;
        float v[8];
        if (ok) load_octet_f32(a.w, a.w_dtype, oct, v);
        else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        float amax = 0.f, vmin = v[0], vmax = v[0];
#pragma unroll
        for (int i = 0; i < 8; ++i) { amax = fmaxf(amax, fabsf(v[i])); vmin = fminf(vmin, v[i]); vmax = fmaxf(vmax, v[i]); }
        for (int m = 1; m < T; m <<= 1) {
            amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, m));
            vmin = fminf(vmin, __shfl_xor_sync(0xFFFFFFFFu, vmin, m));
            vmax = fmaxf(vmax, __shfl_xor_sync(0xFFFFFFFFu, vmax, m));
        }
        if (!ok) continue;
        float scale, zero;
        group_scale(a, amax, vmin, vmax, scale, zero);
        quantise_store<BITS>(a, oct, v, scale, zero);
        if ((oct & (T - 1)) == 0) {
            a.scale[oct / T] = scale;
            if (a.zp != nullptr) a.zp[oct / T] = zero;
        }
    }
}

// Row-wise scales and groups wider than 256 weights: one CTA per group, statistics pass + quantise pass (second read from L2).
template <int BITS>
__global__ void __launch_bounds__(kThreads) weight_quant_wide_kernel(const WQArgs a) {
    __shared__ float s_amax[kThreads / 32], s_min[kThreads / 32], s_max[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t groups = a.octets / a.oct_per_group;
    for (int64_t grp = blockIdx.x; grp < groups; grp += gridDim.x) {
        const int64_t oct0 = grp * a.oct_per_group;
        float amax = 0.f, vmin = INFINITY, vmax = -INFINITY;
        for (int o = threadIdx.x; o < a.oct_per_group; o += kThreads) {
            float v[8];
            load_octet_f32(a.w, a.w_dtype, oct0 + o, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) { amax = fmaxf(amax, fabsf(v[i])); vmin = fminf(vmin, v[i]); vmax = fmaxf(vmax, v[i]); }
        }
        amax = warp_max(amax);
        vmin = warp_min(vmin);
        vmax = warp_max(vmax);
        __syncthreads();                                        // the previous group's readers are done with the exchange buffers
        if (lane == 0) { s_amax[warp] = amax; s_min[warp] = vmin; s_max[warp] = vmax; }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kThreads / 32; ++i) { amax = fmaxf(amax, s_amax[i]); vmin = fminf(vmin, s_min[i]); vmax = fmaxf(vmax, s_max[i]); }
        float scale, zero;
        group_scale(a, amax, vmin, vmax, scale, zero);
        for (int o = threadIdx.x; o < a.oct_per_group; o += kThreads) {
            float v[8];
            load_octet_f32(a.w, a.w_dtype, oct0 + o, v);
            quantise_store<BITS>(a, oct0 + o, v, scale, zero);
        }
        if (threadIdx.x == 0) {
            a.scale[grp] = scale;
            if (a.zp != nullptr) a.zp[grp] = zero;
        }
    }
}

}  // namespace
}  // namespace sdnq

using namespace sdnq;

extern "C" int sdnq_b200_quantize_weight(const void* w, int w_dtype, int64_t N, int64_t K, int64_t group_size, const sdnq_weight_format* fmt,
                                         int scale_dtype, void* codes, float* scale, float* zero_point, void* stream) {
    WFormat f;
    int rc = make_wformat(fmt, &f);
    if (rc != SDNQ_OK) return rc;
    SDNQ_REQUIRE(w && codes && scale, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(f.word_bytes == 1 && f.bits >= 2 && f.bits <= 8, SDNQ_EUNSUPPORTED, "quantize_weight: formats of 2..8 bits (got kind %d, %d bits)", f.kind, f.bits);
    const bool is_float = f.kind != SDNQ_W_INT;
    if (is_float) SDNQ_REQUIRE(f.exponent >= 1 && f.exponent <= 8 && f.mantissa >= 0 && f.mantissa <= 15, SDNQ_EINVAL, "quantize_weight: bad minifloat e%dm%d", f.exponent, f.mantissa);
    SDNQ_REQUIRE(w_dtype == SDNQ_F32 || w_dtype == SDNQ_BF16 || w_dtype == SDNQ_F16, SDNQ_EINVAL, "bad weight dtype %d", w_dtype);
    SDNQ_REQUIRE(scale_dtype == SDNQ_F32 || scale_dtype == SDNQ_BF16 || scale_dtype == SDNQ_F16, SDNQ_EINVAL, "bad scale dtype %d", scale_dtype);
    SDNQ_REQUIRE(N > 0 && K > 0 && K % 8 == 0, SDNQ_EUNSUPPORTED, "quantize_weight: K (=%lld) must be a positive multiple of 8", (long long)K);
    if (group_size <= 0 || group_size > K) group_size = K;
    SDNQ_REQUIRE(K % group_size == 0 && group_size % 8 == 0, SDNQ_EUNSUPPORTED, "quantize_weight: the group size (%lld) must divide K and be a multiple of 8", (long long)group_size);
    SDNQ_REQUIRE(!f.is_unsigned || zero_point != nullptr, SDNQ_EINVAL, "unsigned formats need a zero_point output");
    SDNQ_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(codes) & 7) == 0, SDNQ_EINVAL, "w must be 16-byte and codes 8-byte aligned");
    const int64_t octets = N * K / 8;
    float qmax = f.is_unsigned ? static_cast<float>((1 << f.bits) - 1) : static_cast<float>((1 << (f.bits - 1)) - 1);
    float qmin = f.is_unsigned ? 0.f : -static_cast<float>(1 << (f.bits - 1));
    if (f.kind == SDNQ_W_FP8_E4M3FN) { qmax = 448.f; qmin = -448.f; }
    else if (f.kind == SDNQ_W_FP8_E5M2) { qmax = 57344.f; qmin = -57344.f; }
    else if (is_float) {
        // "fn" minifloats use every exponent code (common.py:16-334): max = (2 - 2^-M) * 2^(2^E - 1 - bias), bias = 2^(E-1) - 1
        const int bias = (1 << (f.exponent - 1)) - 1;
        qmax = std::ldexp(2.0f - std::ldexp(1.0f, -f.mantissa), (1 << f.exponent) - 1 - bias);
        qmin = f.is_unsigned ? 0.f : -qmax;
    }
    WQArgs a{w, w_dtype, octets, static_cast<int>(group_size / 8), f.bits, f.is_unsigned, (f.bits < 8 && !is_float) ? 1 : 0, f.kind, f.exponent, f.mantissa,
             qmin, qmax, scale_dtype, reinterpret_cast<uint8_t*>(codes), scale, f.is_unsigned ? zero_point : nullptr};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int T = a.oct_per_group;
    const bool narrow = T <= 32 && (T & (T - 1)) == 0;
    cudaError_t e = cudaSuccess;
    if (narrow) {
        const int64_t blocks = (octets + kThreads - 1) / kThreads, cap = int64_t(num_sms()) * 16;
        const unsigned grid = static_cast<unsigned>(blocks < cap ? blocks : cap);
        SDNQ_DISPATCH_BITS(f.bits, (e = launch_pdl(weight_quant_group_kernel<BITS>, dim3(grid), dim3(kThreads), 0, st, a)));
    } else {
        const int64_t groups = octets / T, cap = int64_t(num_sms()) * 16;
        const unsigned grid = static_cast<unsigned>(groups < cap ? groups : cap);
        SDNQ_DISPATCH_BITS(f.bits, (e = launch_pdl(weight_quant_wide_kernel<BITS>, dim3(grid), dim3(kThreads), 0, st, a)));
    }
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of weight_quant_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("weight_quant_kernel");
}
