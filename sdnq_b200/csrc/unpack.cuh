// Device-side decoders for SDNQ weight storage formats.
//
// Bit layouts follow reference packed_int/pack.py:201-321 / unpack.py:233-372 (sub-byte integers) and
// packed_float.py:85-132 (eXmY "fn"/"fnu" minifloats); both were restated from the layout block L1 of
// SURVEY.md section 8 and are pinned bit-exactly by tests/ against reference-generated fixtures.
//
// The unit of work everywhere is one *octet*: 8 consecutive logical values = `bits` consecutive
// storage bytes (for 6-bit: two 3-byte groups of 4 values).  A lane that owns an octet reads `bits`
// bytes and produces 8 codes, so a warp reads 32*bits contiguous bytes and emits 256 contiguous values.
#pragma once

#include "common.cuh"

namespace sdnq {

struct WFormat {
    int kind;         // sdnq_wkind
    int bits;         // 1..8
    int is_unsigned;
    int exponent;
    int mantissa;
    int word_bytes;   // 1 or 8 (uint1 stored as one int64 per packed byte)
    int int_offset;   // value added to the unsigned code of a signed integer format (= dtype min)
};

inline int make_wformat(const sdnq_weight_format* f, WFormat* out) {
    SDNQ_REQUIRE(f != nullptr, SDNQ_EINVAL, "weight format is NULL");
    SDNQ_REQUIRE(f->bits >= 1 && f->bits <= 8, SDNQ_EUNSUPPORTED,
                 "weight formats wider than 8 bits are not implemented by the CUDA kernels (got %d bits)", f->bits);
    SDNQ_REQUIRE(f->kind >= SDNQ_W_INT && f->kind <= SDNQ_W_FP8_E5M2, SDNQ_EINVAL, "bad weight kind %d", f->kind);
    SDNQ_REQUIRE(f->word_bytes == 1 || (f->word_bytes == 8 && f->bits == 1), SDNQ_EINVAL,
                 "word_bytes must be 1 (or 8 for 1-bit formats), got %d", f->word_bytes);
    if (f->kind == SDNQ_W_FP8_E4M3FN || f->kind == SDNQ_W_FP8_E5M2)
        SDNQ_REQUIRE(f->bits == 8, SDNQ_EINVAL, "native fp8 formats are 8 bits");
    if (f->kind == SDNQ_W_MINIFLOAT) {
        const int sign = f->is_unsigned ? 0 : 1;
        SDNQ_REQUIRE(f->exponent >= 1 && f->exponent <= 5 && f->mantissa >= 0 && sign + f->exponent + f->mantissa == f->bits,
                     SDNQ_EINVAL, "inconsistent minifloat format: bits=%d e=%d m=%d unsigned=%d", f->bits, f->exponent,
                     f->mantissa, f->is_unsigned);
    }
    out->kind = f->kind;
    out->bits = f->bits;
    out->is_unsigned = f->is_unsigned;
    out->exponent = f->exponent;
    out->mantissa = f->mantissa;
    out->word_bytes = f->word_bytes;
    // packed signed integers are offset-binary; a native int8 byte is plain two's complement
    out->int_offset = (f->kind == SDNQ_W_INT && !f->is_unsigned && f->bits < 8) ? -(1 << (f->bits - 1)) : 0;
    return SDNQ_OK;
}

// ---- read the `BITS` storage bytes of octet `oct` (octets are numbered along the flattened tensor)
template <int BITS>
__device__ __forceinline__ void load_octet_bytes(const uint8_t* __restrict__ base, int64_t oct, int word_bytes,
                                                 uint32_t (&b)[BITS]) {
    if constexpr (BITS == 8) {
        const uint2 r = *reinterpret_cast<const uint2*>(base + oct * 8);
        b[0] = r.x & 0xFF; b[1] = (r.x >> 8) & 0xFF; b[2] = (r.x >> 16) & 0xFF; b[3] = r.x >> 24;
        b[4] = r.y & 0xFF; b[5] = (r.y >> 8) & 0xFF; b[6] = (r.y >> 16) & 0xFF; b[7] = r.y >> 24;
    } else if constexpr (BITS == 4) {
        const uint32_t r = *reinterpret_cast<const uint32_t*>(base + oct * 4);
        b[0] = r & 0xFF; b[1] = (r >> 8) & 0xFF; b[2] = (r >> 16) & 0xFF; b[3] = r >> 24;
    } else if constexpr (BITS == 2) {
        const uint16_t r = *reinterpret_cast<const uint16_t*>(base + oct * 2);
        b[0] = r & 0xFF; b[1] = r >> 8;
    } else if constexpr (BITS == 1) {
        b[0] = base[oct * word_bytes];   // little-endian low byte of the int64 word when word_bytes == 8
    } else {
        const uint8_t* p = base + oct * BITS;
#pragma unroll
        for (int i = 0; i < BITS; ++i) b[i] = p[i];
    }
}

// ---- storage bytes of one octet -> 8 unsigned codes
template <int BITS>
__device__ __forceinline__ void decode_octet(const uint32_t (&b)[BITS], uint32_t (&v)[8]) {
    if constexpr (BITS == 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = b[i];
    } else if constexpr (BITS == 7) {   // b_i = v_i | ((v7 << (i+1)) & 0x80)
        uint32_t top = 0;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            v[i] = b[i] & 0x7F;
            top |= (b[i] >> 7) << (6 - i);
        }
        v[7] = top;
    } else if constexpr (BITS == 6) {   // two groups: 4 values <- 3 bytes, b_i = v_i | ((v3 << 2(i+1)) & 0xC0)
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const uint32_t b0 = b[3 * g], b1 = b[3 * g + 1], b2 = b[3 * g + 2];
            v[4 * g] = b0 & 0x3F;
            v[4 * g + 1] = b1 & 0x3F;
            v[4 * g + 2] = b2 & 0x3F;
            v[4 * g + 3] = ((b0 >> 6) << 4) | ((b1 >> 6) << 2) | (b2 >> 6);
        }
    } else if constexpr (BITS == 5) {
#pragma unroll
        for (int i = 0; i < 5; ++i) v[i] = b[i] & 0x1F;
        v[5] = (b[0] >> 5) | (((b[3] >> 5) & 3) << 3);
        v[6] = (b[1] >> 5) | (((b[4] >> 5) & 3) << 3);
        v[7] = (b[2] >> 5) | ((b[4] >> 7) << 3) | ((b[3] >> 7) << 4);
    } else if constexpr (BITS == 4) {   // 4 bytes, two values each: b = v0 | v1 << 4
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = b[i] & 0xF;
            v[2 * i + 1] = b[i] >> 4;
        }
    } else if constexpr (BITS == 3) {
        v[0] = b[0] & 7; v[1] = b[1] & 7; v[2] = b[2] & 7;
        v[3] = (b[0] >> 3) & 7; v[4] = (b[1] >> 3) & 7; v[5] = (b[2] >> 3) & 7;
        v[6] = (b[0] >> 6) | (((b[2] >> 6) & 1) << 2);
        v[7] = (b[1] >> 6) | ((b[2] >> 7) << 2);
    } else if constexpr (BITS == 2) {   // 2 bytes, four values each
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[4 * i + j] = (b[i] >> (2 * j)) & 3;
        }
    } else {                            // 1 bit: one byte, value i in bit i
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (b[0] >> j) & 1;
    }
}

// ---- eXmY fn/fnu code -> float (every code finite; "-0" decodes to +0 as upstream does)
__device__ __forceinline__ float decode_minifloat(uint32_t c, int bits, int E, int M, int is_unsigned) {
    const uint32_t magbits = is_unsigned ? bits : bits - 1;
    const uint32_t mag = c & ((1u << magbits) - 1u);
    const uint32_t sign = is_unsigned ? 0u : (c >> (bits - 1)) & 1u;
    const uint32_t e = mag >> M;
    const uint32_t m = mag & ((1u << M) - 1u);
    const int bias = (1 << (E - 1)) - 1;
    float val;
    if (e == 0) {
        val = static_cast<float>(m) * __uint_as_float(static_cast<uint32_t>(1 - bias - M + 127) << 23);  // subnormal, exact
    } else {
        // (1 + m/2^M) * 2^(e-bias): build the fp32 pattern directly
        const uint32_t f32 = ((e - bias + 127u) << 23) | (m << (23 - M));
        val = __uint_as_float(f32);
    }
    return (sign && mag != 0u) ? -val : val;
}

// ---- one octet -> 8 real-valued codes (before scale): integer code (+ signed offset) or float value
template <int BITS>
__device__ __forceinline__ void octet_values(const uint8_t* __restrict__ base, int64_t oct, const WFormat& f,
                                             float (&q)[8], uint32_t (&codes)[8]) {
    uint32_t b[BITS];
    load_octet_bytes<BITS>(base, oct, f.word_bytes, b);
    decode_octet<BITS>(b, codes);
    if (f.kind == SDNQ_W_INT) {
        if (BITS == 8 && !f.is_unsigned) {
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = static_cast<float>(static_cast<int8_t>(codes[i]));
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = static_cast<float>(static_cast<int>(codes[i]) + f.int_offset);
        }
    } else if (f.kind == SDNQ_W_MINIFLOAT) {
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = decode_minifloat(codes[i], f.bits, f.exponent, f.mantissa, f.is_unsigned);
    } else if (f.kind == SDNQ_W_FP8_E4M3FN) {
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = e4m3_to_f32(static_cast<uint8_t>(codes[i]));
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = e5m2_to_f32(static_cast<uint8_t>(codes[i]));
    }
}

#define SDNQ_DISPATCH_BITS(bits, ...)                       \
    switch (bits) {                                         \
        case 1: { constexpr int BITS = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int BITS = 2; __VA_ARGS__; } break; \
        case 3: { constexpr int BITS = 3; __VA_ARGS__; } break; \
        case 4: { constexpr int BITS = 4; __VA_ARGS__; } break; \
        case 5: { constexpr int BITS = 5; __VA_ARGS__; } break; \
        case 6: { constexpr int BITS = 6; __VA_ARGS__; } break; \
        case 7: { constexpr int BITS = 7; __VA_ARGS__; } break; \
        default: { constexpr int BITS = 8; __VA_ARGS__; } break; \
    }

}  // namespace sdnq
