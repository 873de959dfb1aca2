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

// ---- read the `BITS` storage bytes of octet `oct` (octets are numbered along the flattened tensor) as packed
// little-endian 32-bit words (byte i of the octet = bits [8i, 8i+8) of the word array); widest aligned loads possible
template <int BITS> struct OctetWords { static constexpr int N = (BITS + 3) / 4; };

template <int BITS>
__device__ __forceinline__ void load_octet_bytes(const uint8_t* __restrict__ base, int64_t oct, int word_bytes,
                                                 uint32_t (&w)[OctetWords<BITS>::N]) {
    if constexpr (BITS == 8) {
        const uint2 r = *reinterpret_cast<const uint2*>(base + oct * 8);
        w[0] = r.x; w[1] = r.y;
    } else if constexpr (BITS == 4) {
        w[0] = *reinterpret_cast<const uint32_t*>(base + oct * 4);
    } else if constexpr (BITS == 2) {
        w[0] = *reinterpret_cast<const uint16_t*>(base + oct * 2);
    } else if constexpr (BITS == 1) {
        w[0] = base[oct * word_bytes];   // little-endian low byte of the int64 word when word_bytes == 8
    } else if constexpr (BITS == 6) {    // 6 bytes, 2-byte aligned
        const uint16_t* p = reinterpret_cast<const uint16_t*>(base + oct * 6);
        w[0] = uint32_t(p[0]) | (uint32_t(p[1]) << 16);
        w[1] = p[2];
    } else {                             // 3, 5, 7 bytes at an odd offset: byte loads
        const uint8_t* p = base + oct * BITS;
#pragma unroll
        for (int i = 0; i < OctetWords<BITS>::N; ++i) w[i] = 0;
#pragma unroll
        for (int i = 0; i < BITS; ++i) w[i >> 2] |= uint32_t(p[i]) << (8 * (i & 3));
    }
}

template <int NW>
__device__ __forceinline__ uint32_t octet_byte(const uint32_t (&w)[NW], int i) { return (w[i >> 2] >> (8 * (i & 3))) & 0xFFu; }

// ---- storage bytes of one octet -> 8 unsigned codes
template <int BITS>
__device__ __forceinline__ void decode_octet(const uint32_t (&w)[OctetWords<BITS>::N], uint32_t (&v)[8]) {
    if constexpr (BITS == 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = octet_byte(w, i);
    } else if constexpr (BITS == 7) {   // b_i = v_i | ((v7 << (i+1)) & 0x80)
        uint32_t top = 0;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const uint32_t b = octet_byte(w, i);
            v[i] = b & 0x7F;
            top |= (b >> 7) << (6 - i);
        }
        v[7] = top;
    } else if constexpr (BITS == 6) {   // two groups: 4 values <- 3 bytes, b_i = v_i | ((v3 << 2(i+1)) & 0xC0)
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const uint32_t b0 = octet_byte(w, 3 * g), b1 = octet_byte(w, 3 * g + 1), b2 = octet_byte(w, 3 * g + 2);
            v[4 * g] = b0 & 0x3F;
            v[4 * g + 1] = b1 & 0x3F;
            v[4 * g + 2] = b2 & 0x3F;
            v[4 * g + 3] = ((b0 >> 6) << 4) | ((b1 >> 6) << 2) | (b2 >> 6);
        }
    } else if constexpr (BITS == 5) {
        const uint32_t b0 = octet_byte(w, 0), b1 = octet_byte(w, 1), b2 = octet_byte(w, 2), b3 = octet_byte(w, 3), b4 = octet_byte(w, 4);
        v[0] = b0 & 0x1F; v[1] = b1 & 0x1F; v[2] = b2 & 0x1F; v[3] = b3 & 0x1F; v[4] = b4 & 0x1F;
        v[5] = (b0 >> 5) | (((b3 >> 5) & 3) << 3);
        v[6] = (b1 >> 5) | (((b4 >> 5) & 3) << 3);
        v[7] = (b2 >> 5) | ((b4 >> 7) << 3) | ((b3 >> 7) << 4);
    } else if constexpr (BITS == 4) {   // 4 bytes, two values each: b = v0 | v1 << 4  ==  nibble i of the word
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (w[0] >> (4 * i)) & 0xF;
    } else if constexpr (BITS == 3) {
        const uint32_t b0 = octet_byte(w, 0), b1 = octet_byte(w, 1), b2 = octet_byte(w, 2);
        v[0] = b0 & 7; v[1] = b1 & 7; v[2] = b2 & 7;
        v[3] = (b0 >> 3) & 7; v[4] = (b1 >> 3) & 7; v[5] = (b2 >> 3) & 7;
        v[6] = (b0 >> 6) | (((b2 >> 6) & 1) << 2);
        v[7] = (b1 >> 6) | ((b2 >> 7) << 2);
    } else if constexpr (BITS == 2) {   // 2 bytes, four values each  ==  2-bit field i of the 16-bit word
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (w[0] >> (2 * i)) & 3;
    } else {                            // 1 bit: one byte, value i in bit i
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (w[0] >> j) & 1;
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

// ---- 8 codes -> 8 real values (before scale): integer code (+ signed offset) or minifloat / fp8 value
template <int BITS>
__device__ __forceinline__ void codes_to_values(const uint32_t (&codes)[8], const WFormat& f, float (&q)[8]) {
    if (f.kind == SDNQ_W_INT) {
        // exact int -> float without the conversion pipe: 0x4B000000 | c is the float 2^23 + c for c < 2^23
        if (BITS == 8 && !f.is_unsigned) {
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = __uint_as_float(0x4B000000u | (codes[i] ^ 0x80u)) - 8388736.0f;      // two's complement byte
        } else {
            const float bias = 8388608.0f - static_cast<float>(f.int_offset);
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = __uint_as_float(0x4B000000u | codes[i]) - bias;
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

// ---- one octet of storage -> 8 real values
template <int BITS>
__device__ __forceinline__ void octet_values(const uint8_t* __restrict__ base, int64_t oct, const WFormat& f,
                                             float (&q)[8], uint32_t (&codes)[8]) {
    uint32_t w[OctetWords<BITS>::N];
    load_octet_bytes<BITS>(base, oct, f.word_bytes, w);
    decode_octet<BITS>(w, codes);
    codes_to_values<BITS>(codes, f, q);
}

// ---- integer octet -> 8 floats (code + signed offset) through the magic-number trick: byte b -> 0x4B0000bb = 2^23 + b.
// `bias` = 2^23 - offset (2^23 + 128 for two's-complement int8 with flip = 0x80808080); int4 / int8 use one PRMT per element.
template <int BITS>
__device__ __forceinline__ void octet_to_floats(const uint32_t (&w)[OctetWords<BITS>::N], uint32_t flip, float bias, float (&q)[8]) {
    if constexpr (BITS == 4) {
        const uint32_t lo = w[0] & 0x0F0F0F0Fu, hi = (w[0] >> 4) & 0x0F0F0F0Fu;       // even / odd nibbles, one per byte
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            q[2 * i] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7440 | i)) - bias;
            q[2 * i + 1] = __uint_as_float(__byte_perm(hi, 0x4B000000u, 0x7440 | i)) - bias;
        }
    } else if constexpr (BITS == 8) {
        const uint32_t a = w[0] ^ flip, b = w[1] ^ flip;                              // flip = 0x80808080 for two's complement bytes
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            q[i] = __uint_as_float(__byte_perm(a, 0x4B000000u, 0x7440 | i)) - bias;
            q[4 + i] = __uint_as_float(__byte_perm(b, 0x4B000000u, 0x7440 | i)) - bias;
        }
    } else {
        uint32_t codes[8];
        decode_octet<BITS>(w, codes);
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = __uint_as_float(0x4B000000u | codes[i]) - bias;
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
