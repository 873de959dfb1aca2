#include <cstdlib>
// Host stand-ins for the CUDA device intrinsics used by the per-lane arithmetic in sdnq_b200/csrc/{common,unpack}.cuh, so that
// g++ can compile those headers unchanged and tests/test_device_arithmetic_on_host.py can check the decode / convert
// functions bit-for-bit against the oracle without a GPU.  Test infrastructure only; warp-collective code (shuffles) is
// declared but never executed here.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

// every CUDA header the kernels include, up front, so that __shared__ can be redefined once after them: kernels declare shared
// memory as function-local __shared__ arrays; on the host that is one static array per kernel instantiation, shared by the
// lock-stepped threads of the (single) running CTA
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <vector_types.h>
#include <vector_functions.h>
#undef __shared__
#define __shared__ static

#define SDNQ_HOST_EMU 1
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#include "warp.h"

static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
// PRMT in its default mode: result byte i = byte (selector nibble i & 7) of the 8-byte pool {y:x}; nibble bit 3 replicates the sign
[[noreturn]] static inline void __trap() { std::abort(); }      // a device trap kills the context; on the host: abort
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t pool = (uint64_t(y) << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t b = uint32_t(pool >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) b = (b & 0x80) ? 0xFF : 0x00;
        r |= b << (8 * i);
    }
    return r;
}
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsub_rn(float a, float b) { return a - b; }          // built with -ffp-contract=off: no fusion on the host either
static inline float __fmul_rn(float a, float b) { return a * b; }
// cvt.rni.s32.f32: round to nearest even, NaN -> 0, saturating
static inline int __float2int_rn(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return -2147483647 - 1;
    return static_cast<int>(std::nearbyint(v));
}
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline unsigned int __umulhi(unsigned int a, unsigned int b) { return static_cast<unsigned int>((static_cast<unsigned long long>(a) * b) >> 32); }
// the lock-stepped lanes are real threads: atomics must be atomic
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned int atomicMax(unsigned int* p, unsigned int v) {
    unsigned int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz(static_cast<unsigned>(v)); }
// xor-shuffle across the 32 lock-stepped host lanes (warp.h); outside run_warp() nothing calls it
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    const int lane = sdnq_emu::lane_id();
    std::memcpy(&sdnq_emu::xchg()[lane][0], &v, 4);
    sdnq_emu::warp_sync();
    T r;
    std::memcpy(&r, &sdnq_emu::xchg()[lane ^ m][0], 4);
    sdnq_emu::warp_sync();
    return r;
}
// threadIdx.x / blockIdx.x / gridDim.x of the calling host thread (warp.h: run_grid)
struct EmuThreadIdx { struct X { operator int() const { return sdnq_emu::t_tid; } } x; };
struct EmuBlockIdx {
    struct X { operator int() const { return sdnq_emu::g_block.x; } } x;
    struct Y { operator int() const { return sdnq_emu::g_block.y; } } y;
    struct Z { operator int() const { return sdnq_emu::g_block.z; } } z;
};
struct EmuGridDim {
    struct X { operator int() const { return sdnq_emu::g_grid.x; } } x;
    struct Y { operator int() const { return sdnq_emu::g_grid.y; } } y;
    struct Z { operator int() const { return sdnq_emu::g_grid.z; } } z;
};
static const EmuThreadIdx threadIdx = {};
static const EmuBlockIdx blockIdx = {};
static const EmuGridDim gridDim = {};
static inline void __syncthreads() { sdnq_emu::cta_sync(); }
