// Hadamard rotation of a 256-element chunk on the tensor cores (mma.sync m16n8k16, f32 accumulate).
//
// Reference behaviour: rotate_hadamard (quant_utils.py:193-209) = x.unflatten(-1, (-1, G)) @ H_G as a matmul in x.dtype
// (f32 accumulate, result rounded to x.dtype), H_G from build_hadamard_n4 / _n2 (quant_utils.py:144-165).
//
// The butterfly version of this (hadamard_warp in common.cuh) is issue-bound: 5 cross-lane stages = 40 SHFL + the sign /
// permutation bookkeeping of the H4 family come to ~300 warp-instructions per chunk (ncu: 80 % issue-active, 2.2 TB/s).
// Here the chunk is viewed as a 16x16 matrix X (element 16 r + c) and both Hadamard families factor as Kronecker products
// over (r, c):   rotated = L^T . X . R   with 16x16 sign matrices
//     G >= 16 :  R = H_16 of the family,  L = I_(256/G) (x) H_(G/16)        (block diagonal over the groups of the chunk)
//     G <  16 :  R = I_(16/G) (x) H_G,    L = I                               (second product skipped)
//   MMA 1:  T = X . R      A = X straight from the two coalesced 8-byte loads of the lane (lane l owns elements [4l, 4l+4)
//           and [128+4l, 128+4l+4) of the chunk -- the k-slots and the output columns of the fragment are permuted so that
//           this is exactly the A / C fragment), B = per-lane constants (+-1).  Exact: 16-bit inputs, f32 accumulate.
//   T is split into bf16 pieces hi + lo (residual <= 2^-17 relative; a third piece for f16 inputs) and transposed across the warp with 8
//   movmatrix (the C fragment has r on the lane-group axis, the contraction needs it on the k axis).
//   MMA 2:  Y = L^T . T    A = per-lane constants, B = transposed T (hi, then lo accumulated on top).
// The result comes back in the same places the inputs were loaded from, so stores are the same two coalesced accesses.
// 6 HMMA + 8 MOVM + ~24 ALU per chunk instead of ~200 for the butterflies; no shuffles.
#pragma once
#include "common.cuh"

// SDNQ_HOST_EMU (tests/host_emu only): the three warp primitives below (cvt pack, mma.sync, movmatrix) are replaced by host
// models running 32 lock-stepped threads, so tests/test_device_arithmetic_on_host.py can execute Rotation<T>::apply on a CPU.

namespace sdnq {
namespace hadtc {

// Entries of the sign matrices (G = 2^lg).  Family of H_n: n a power of 4 -> kron^k(H4) with H4[a][b] = -1 iff a + b == 3
// (<=> a ^ b == 3 for 2-bit digits); otherwise Sylvester, H[i][j] = (-1)^popc(i & j).  `mask` = n - 1 selects the index bits
// that belong to H_n.  All of this is evaluated at compile time into the fragment table below.
constexpr int popc8(int v) { int n = 0; for (int i = 0; i < 8; ++i) n += (v >> i) & 1; return n; }
constexpr int hsign(bool pow4, int mask, int i, int j) {
    const int z = i ^ j;
    const int odd = pow4 ? popc8(z & (z >> 1) & 0x55 & mask) : popc8(i & j & mask);
    return (odd & 1) ? -1 : 1;
}
// R[c][c2] and L[r][r2] of a chunk rotated in groups of G = 2^lg (0 where the block-diagonal structure has no entry)
constexpr int right_entry(int lg, int c, int c2) {
    const bool pow4 = (lg & 1) == 0;
    if (lg >= 4) return hsign(pow4, 15, c, c2);
    return ((c ^ c2) >> lg) ? 0 : hsign(pow4, (1 << lg) - 1, c, c2);
}
constexpr int left_entry(int lg, int r, int r2) {       // lg > 4: L = I (x) H_(G/16)
    const int lgl = lg - 4;
    return ((r ^ r2) >> lgl) ? 0 : hsign((lgl & 1) == 0, (1 << lgl) - 1, r, r2);
}
// fragment slot -> column of X: lane t's k-slots {2t, 2t+1, 2t+8, 2t+9} are the four consecutive columns 4t .. 4t+3
constexpr int slot_col(int s) { return 4 * ((s & 7) >> 1) + (s & 1) + 2 * (s >> 3); }

// Per-lane constant fragments for every group size (lg = 2..8) and both 16-bit formats of MMA 1 (MMA 2 is always bf16):
// words 0-3 = B fragments of MMA 1 ([n-block][reg]), words 4-7 = A fragment of MMA 2.  For the H4 family (lg even) the
// normalisation 1/sqrt(G) = 2^-(lg/2) is exact in every format, so it is folded into the constants of the last product
// (exact: a power-of-two scale commutes with every rounding here) and the rotation needs no multiply afterwards.
struct FragTable { uint32_t w[7][2][32][8]; };
// +-2^-shift as f16 / bf16 bits (0 for sign == 0)
constexpr uint32_t half_bits(int sign, bool f16, int shift = 0) {
    return sign == 0 ? 0u : (sign < 0 ? 0x8000u : 0u) | (f16 ? uint32_t(15 - shift) << 10 : uint32_t(127 - shift) << 7);
}
constexpr FragTable make_frag_table() {
    FragTable t{};
    for (int lg = 2; lg <= 8; ++lg)
        for (int f = 0; f < 2; ++f)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, q = lane & 3;
                uint32_t* w = t.w[lg - 2][f][lane];
                const int fold = (lg & 1) == 0 ? lg / 2 : 0;          // folded into MMA 1 when it is the only product, else MMA 2
                const int s1 = lg > 4 ? 0 : fold, s2 = lg > 4 ? fold : 0;
                for (int j = 0; j < 2; ++j) {
                    const int n = slot_col(8 * j + g);
                    w[2 * j] = half_bits(right_entry(lg, slot_col(2 * q), n), f, s1) | (half_bits(right_entry(lg, slot_col(2 * q + 1), n), f, s1) << 16);
                    w[2 * j + 1] = half_bits(right_entry(lg, slot_col(2 * q + 8), n), f, s1) | (half_bits(right_entry(lg, slot_col(2 * q + 9), n), f, s1) << 16);
                }
                if (lg > 4) {   // A2[r'][r] = L[r][r']:  a0 = (row g, k 2q..), a1 = (row g+8, k 2q..), a2 = (row g, k 2q+8..), a3 = (row g+8, k 2q+8..)
                    w[4] = half_bits(left_entry(lg, 2 * q, g), false, s2) | (half_bits(left_entry(lg, 2 * q + 1, g), false, s2) << 16);
                    w[5] = half_bits(left_entry(lg, 2 * q, g + 8), false, s2) | (half_bits(left_entry(lg, 2 * q + 1, g + 8), false, s2) << 16);
                    w[6] = half_bits(left_entry(lg, 2 * q + 8, g), false, s2) | (half_bits(left_entry(lg, 2 * q + 9, g), false, s2) << 16);
                    w[7] = half_bits(left_entry(lg, 2 * q + 8, g + 8), false, s2) | (half_bits(left_entry(lg, 2 * q + 9, g + 8), false, s2) << 16);
                }
            }
    return t;
}
__device__ const FragTable g_frag_table = make_frag_table();

template <typename T> struct Half16;
template <> struct Half16<__nv_bfloat16> {
    static constexpr uint32_t kOne = 0x3F80u, kMinusOne = 0xBF80u;
    __device__ static __forceinline__ uint32_t pack(float lo, float hi) {
#ifdef SDNQ_HOST_EMU
        return ::sdnq_emu::pack16x2(false, lo, hi);
#else
        uint32_t r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
#endif
    }
    __device__ static __forceinline__ float lo(uint32_t w) { return __uint_as_float(w << 16); }
    __device__ static __forceinline__ float hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
    __device__ static __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
#ifdef SDNQ_HOST_EMU
        ::sdnq_emu::mma_m16n8k16(false, d, a0, a1, a2, a3, b0, b1);
#else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
#endif
    }
};
template <> struct Half16<__half> {
    static constexpr uint32_t kOne = 0x3C00u, kMinusOne = 0xBC00u;
    __device__ static __forceinline__ uint32_t pack(float lo, float hi) {
#ifdef SDNQ_HOST_EMU
        return ::sdnq_emu::pack16x2(true, lo, hi);
#else
        uint32_t r;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
#endif
    }
    __device__ static __forceinline__ float lo(uint32_t w) { return __half2float(__ushort_as_half(static_cast<unsigned short>(w & 0xFFFFu))); }
    __device__ static __forceinline__ float hi(uint32_t w) { return __half2float(__ushort_as_half(static_cast<unsigned short>(w >> 16))); }
    __device__ static __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
#ifdef SDNQ_HOST_EMU
        ::sdnq_emu::mma_m16n8k16(true, d, a0, a1, a2, a3, b0, b1);
#else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
#endif
    }
};

__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
#ifdef SDNQ_HOST_EMU
    return ::sdnq_emu::movmatrix_trans(a);
#else
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
#endif
}

// Per-lane constant fragments: two 16-byte loads per thread from the table.
template <typename T>
struct Rotation {
    uint32_t b1[4];      // MMA 1 B fragments: [n-block j][reg]
    uint32_t a2[4];      // MMA 2 A fragment
    bool two_sided;
    bool folded;         // 1/sqrt(G) is already in the constants (H4 family)

    __device__ __forceinline__ void init(int G, int lane) {
        const int lg = 31 - __clz(G);
        two_sided = G > 16;
        folded = (lg & 1) == 0;
        constexpr int f = ElemTraits<T>::kDtype == SDNQ_F16 ? 1 : 0;
        const uint4* w = reinterpret_cast<const uint4*>(g_frag_table.w[lg - 2][f][lane]);
        const uint4 lo = w[0], hi = w[1];
        b1[0] = lo.x; b1[1] = lo.y; b1[2] = lo.z; b1[3] = lo.w;
        a2[0] = hi.x; a2[1] = hi.y; a2[2] = hi.z; a2[3] = hi.w;
    }

    // raw: x = elements (4l, 4l+1), y = (4l+2, 4l+3), z = (128+4l, +1), w = (128+4l+2, +3) of the chunk, as 16-bit pairs.
    // On return the same places hold the rotated chunk times `factor`, rounded to T.  All 32 lanes must call this.
    // Returns max |rotated value| over the lane's 8 outputs *before* the rounding to T: rounding is monotone and symmetric, so
    // round(max |y|) == max |round(y)| and the caller can round the row maximum once instead of re-reading the packed values.
    __device__ __forceinline__ float apply(uint4& raw, float factor) const {
        float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f};
        Half16<T>::mma(t0, raw.x, raw.z, raw.y, raw.w, b1[0], b1[1]);      // columns 4t', 4t'+1     of rows g, g+8
        Half16<T>::mma(t1, raw.x, raw.z, raw.y, raw.w, b1[2], b1[3]);      // columns 4t'+2, 4t'+3
        if (two_sided) {
            float y0[4] = {0.f, 0.f, 0.f, 0.f}, y1[4] = {0.f, 0.f, 0.f, 0.f};
            rotate_rows(t0, y0);
            rotate_rows(t1, y1);
#pragma unroll
            for (int i = 0; i < 4; ++i) { t0[i] = y0[i]; t1[i] = y1[i]; }
        }
        if (!folded) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { t0[i] *= factor; t1[i] *= factor; }
        }
        raw.x = Half16<T>::pack(t0[0], t0[1]);
        raw.y = Half16<T>::pack(t1[0], t1[1]);
        raw.z = Half16<T>::pack(t0[2], t0[3]);
        raw.w = Half16<T>::pack(t1[2], t1[3]);
        return fmaxf(fmaxf(fmaxf(fabsf(t0[0]), fabsf(t0[1])), fmaxf(fabsf(t0[2]), fabsf(t0[3]))),
                     fmaxf(fmaxf(fabsf(t1[0]), fabsf(t1[1])), fmaxf(fabsf(t1[2]), fabsf(t1[3]))));
    }

    // y += L^T . t for one n-block: t = C fragment (rows g / g+8) -> bf16 pieces (hi, lo[, lo2]: 8 bits each) ->
    // movmatrix -> B fragments.  MMA 2 always runs in bf16 (the constants are +-1, the pieces carry f32's exponent range,
    // so f16 activations lose nothing to f16 subnormals); f16 inputs take a third piece so the split error stays far
    // below an f16 ulp.
    __device__ __forceinline__ void rotate_rows(const float (&t)[4], float (&y)[4]) const {
        using B = Half16<__nv_bfloat16>;
        constexpr int kPieces = sizeof(T) == 2 && ElemTraits<T>::kDtype == SDNQ_F16 ? 3 : 2;
        float r[4] = {t[0], t[1], t[2], t[3]};
        uint32_t pa[kPieces], pb[kPieces];
#pragma unroll
        for (int p = 0; p < kPieces; ++p) {
            pa[p] = B::pack(r[0], r[1]);
            pb[p] = B::pack(r[2], r[3]);
            if (p + 1 < kPieces) {
                r[0] -= B::lo(pa[p]); r[1] -= B::hi(pa[p]);
                r[2] -= B::lo(pb[p]); r[3] -= B::hi(pb[p]);
            }
        }
#pragma unroll
        for (int p = kPieces - 1; p >= 0; --p) {                            // smallest piece first
            const uint32_t b0 = movmatrix_trans(pa[p]), b1r = movmatrix_trans(pb[p]);
            B::mma(y, a2[0], a2[1], a2[2], a2[3], b0, b1r);
        }
    }
};

}  // namespace hadtc
}  // namespace sdnq
