// Lock-step warps / CTAs on the host: every CUDA thread is a pthread, warp collectives and __syncthreads meet at barriers.  Host models of the three
// PTX warp primitives hadamard_tc.cuh uses, following the PTX ISA fragment layouts (g = lane >> 2, q = lane & 3):
//   mma.sync.m16n8k16 row.col   A 16x16: a0 = (row g,   k 2q,2q+1)  a1 = (row g+8, k 2q,2q+1)  a2 = (row g, k 2q+8,2q+9)  a3 = (row g+8, k 2q+8,2q+9)
//                               B 16x8 : b0 = (k 2q,2q+1,   col g)  b1 = (k 2q+8,2q+9, col g)
//                               C/D 16x8: d0,d1 = (row g, col 2q,2q+1)   d2,d3 = (row g+8, col 2q,2q+1)
//   movmatrix.m8n8.trans.b16    lane holds (row g, cols 2q,2q+1) of an 8x8 matrix of 16-bit elements; result = its transpose, same layout
//   cvt.rn.{bf16x2,f16x2}.f32   two floats -> packed pair, round to nearest even (low half = first operand `lo`)
// Products of 16-bit values are exact in double; the sum is formed in double and rounded once to f32, which equals the tensor
// core's f32 accumulation whenever the exact sum is representable (the tests use such inputs for their bit-exact checks).
#pragma once
#include <pthread.h>

#include <cmath>

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

namespace sdnq_emu {

constexpr int kMaxWarps = 8;
inline thread_local int t_tid = 0;                    // threadIdx.x of the calling host thread
inline dim3 g_block(0, 0, 0), g_grid(1, 1, 1);        // blockIdx / gridDim (CTAs run one after another)
inline pthread_barrier_t g_warp_barrier[kMaxWarps], g_cta_barrier;
inline uint32_t g_xchg_all[kMaxWarps][32][6];

inline int lane_id() { return t_tid & 31; }
inline uint32_t (*xchg())[6] { return g_xchg_all[t_tid >> 5]; }
inline void warp_sync() { pthread_barrier_wait(&g_warp_barrier[t_tid >> 5]); }
inline void cta_sync() { pthread_barrier_wait(&g_cta_barrier); }

inline float half16_to_float(bool f16, uint32_t bits16) {
    if (f16) {
        __half_raw r; r.x = static_cast<unsigned short>(bits16);
        return __half2float(__half(r));
    }
    uint32_t u = bits16 << 16; float f; std::memcpy(&f, &u, 4); return f;
}

inline uint32_t pack16x2(bool f16, float lo, float hi) {
    if (f16) {
        const __half a = __float2half_rn(lo), b = __float2half_rn(hi);
        return uint32_t(__half_raw(a).x) | (uint32_t(__half_raw(b).x) << 16);
    }
    const __nv_bfloat16 a = __float2bfloat16_rn(lo), b = __float2bfloat16_rn(hi);
    return uint32_t(__nv_bfloat16_raw(a).x) | (uint32_t(__nv_bfloat16_raw(b).x) << 16);
}

inline void mma_m16n8k16(bool f16, float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    const int lane = lane_id(), g = lane >> 2, q = lane & 3;
    uint32_t (*g_xchg)[6] = xchg();
    uint32_t* mine = g_xchg[lane];
    mine[0] = a0; mine[1] = a1; mine[2] = a2; mine[3] = a3; mine[4] = b0; mine[5] = b1;
    warp_sync();
    auto A = [&](int row, int k) {                       // element (row, k) of the 16x16 A operand
        const int src = (row & 7) * 4 + ((k & 7) >> 1);
        const int reg = (row >> 3) + 2 * (k >> 3);
        return half16_to_float(f16, (g_xchg[src][reg] >> (16 * (k & 1))) & 0xFFFFu);
    };
    auto B = [&](int k, int col) {                       // element (k, col) of the 16x8 B operand
        const int src = col * 4 + ((k & 7) >> 1);
        return half16_to_float(f16, (g_xchg[src][4 + (k >> 3)] >> (16 * (k & 1))) & 0xFFFFu);
    };
    for (int i = 0; i < 4; ++i) {
        const int row = g + 8 * (i >> 1), col = 2 * q + (i & 1);
        double acc = d[i];
        for (int k = 0; k < 16; ++k) acc += double(A(row, k)) * double(B(k, col));
        d[i] = static_cast<float>(acc);
    }
    warp_sync();                                         // everyone has read before the exchange area is reused
}

inline uint32_t movmatrix_trans(uint32_t a) {
    const int lane = lane_id(), g = lane >> 2, q = lane & 3;
    uint32_t (*g_xchg)[6] = xchg();
    g_xchg[lane][0] = a;
    warp_sync();
    uint32_t r = 0;
    for (int e = 0; e < 2; ++e) {                        // output element (row g, col 2q+e) = input element (row 2q+e, col g)
        const int row = 2 * q + e, col = g;
        const uint32_t w = g_xchg[row * 4 + (col >> 1)][0];
        r |= ((w >> (16 * (col & 1))) & 0xFFFFu) << (16 * e);
    }
    warp_sync();
    return r;
}

// rcp.approx.ftz.f32: the hardware result is within 1 ulp of 1/x; the host model returns the correctly rounded reciprocal
// (subnormal inputs / outputs flushed to zero).  RowDivider refines it to a correctly rounded quotient either way.
inline float rcp_approx_ftz(float x) {
    if (std::fabs(x) < 1.17549435e-38f) x = std::copysign(0.0f, x);
    float r = 1.0f / x;
    if (std::fabs(r) < 1.17549435e-38f) r = std::copysign(0.0f, r);
    return r;
}
// cvt.rn.f16x2.e4m3x2: two e4m3 bytes -> two f16 (exact), low byte -> low half
inline uint32_t e4m3x2_to_f16x2(unsigned short pair) {
    const __half_raw lo = __nv_cvt_fp8_to_halfraw(static_cast<__nv_fp8_storage_t>(pair & 0xFF), __NV_E4M3);
    const __half_raw hi = __nv_cvt_fp8_to_halfraw(static_cast<__nv_fp8_storage_t>(pair >> 8), __NV_E4M3);
    return uint32_t(lo.x) | (uint32_t(hi.x) << 16);
}
inline uint32_t sat8(int v) { return static_cast<uint32_t>(static_cast<uint8_t>(static_cast<int8_t>(v < -128 ? -128 : (v > 127 ? 127 : v)))); }
// two cvt.pack.sat.s8.s32.b32: bytes (low -> high) = sat8(c0), sat8(c1), sat8(c2), sat8(c3)
inline uint32_t pack_sat_s8x4(int c0, int c1, int c2, int c3) { return sat8(c0) | (sat8(c1) << 8) | (sat8(c2) << 16) | (sat8(c3) << 24); }

// run fn() as a grid of CTAs of `threads` lock-stepped host threads each (CTAs one after another); the device code reads
// threadIdx.x / blockIdx / gridDim through the stand-ins in prelude.h
template <typename F>
void run_grid(dim3 grid, int threads, F&& fn) {
    const int warps = (threads + 31) / 32;
    const unsigned total = grid.x * grid.y * grid.z;
    g_grid = grid;
    pthread_barrier_t between;                            // separates consecutive CTAs (distinct from the kernel's own __syncthreads barrier)
    pthread_barrier_init(&between, nullptr, threads);
    pthread_barrier_init(&g_cta_barrier, nullptr, threads);
    for (int w = 0; w < warps; ++w) pthread_barrier_init(&g_warp_barrier[w], nullptr, threads - 32 * w < 32 ? threads - 32 * w : 32);
    struct Arg { F* fn; int tid; unsigned total; dim3 grid; pthread_barrier_t* between; } args[kMaxWarps * 32];
    pthread_t th[kMaxWarps * 32];
    for (int t = 0; t < threads; ++t) {
        args[t] = {&fn, t, total, grid, &between};
        pthread_create(&th[t], nullptr, [](void* p) -> void* {        // one host thread per CUDA thread, reused for every CTA of the grid
            Arg* a = static_cast<Arg*>(p);
            t_tid = a->tid;
            for (unsigned b = 0; b < a->total; ++b) {
                if (a->tid == 0) g_block = dim3(b % a->grid.x, (b / a->grid.x) % a->grid.y, b / (a->grid.x * a->grid.y));
                pthread_barrier_wait(a->between);
                (*a->fn)();
                pthread_barrier_wait(a->between);
            }
            return nullptr;
        }, &args[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], nullptr);
    pthread_barrier_destroy(&between);
    pthread_barrier_destroy(&g_cta_barrier);
    for (int w = 0; w < warps; ++w) pthread_barrier_destroy(&g_warp_barrier[w]);
}
template <typename F>
void run_grid(int grid, int threads, F&& fn) { run_grid(dim3(grid), threads, static_cast<F&&>(fn)); }

// one warp: fn(lane)
template <typename F>
void run_warp(F&& fn) {
    run_grid(1, 32, [&] { fn(lane_id()); });
}

}  // namespace sdnq_emu
