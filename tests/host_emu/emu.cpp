// g++ build of the per-lane device arithmetic of the sdnq_b200 kernels (see prelude.h).  Exposes a tiny C ABI for the test.
#include "prelude.h"
#include "../../sdnq_b200/csrc/unpack.cuh"

#include <cstdarg>
#include <cstdio>

using namespace sdnq;

// ---- what capi.cu provides in the real library (no CUDA runtime here)
namespace sdnq {
namespace {
thread_local char g_error[512] = "";
int64_t g_launches = 0;
}  // namespace
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}
void count_launch(int n) { g_launches += n; }
int num_sms() { return 2; }                     // grids are sized from this: keep the emulated grids small
bool pdl_enabled() { return true; }
// dequant_svd.cu (tcgen05) is outside the emulator: "not covered" sends SVD layers to the CUDA-core update of dequant.cu
int dequant_svd_tc(const void*, const WFormat&, const float*, const float*, int64_t, int64_t, int, int, int, int, const void*, int64_t, int64_t,
                   const void*, int64_t, int64_t, int, int, void*, int, cudaStream_t) { return 1; }
// batched K3s: same translation unit as dequant_svd_tc -- the plan reports "not covered"
size_t svd_batch_entry_bytes() { return 512; }
int svd_batch_fill(void*, int, int, const void*, const WFormat&, const float*, const float*, int64_t, int64_t, int, int, int, int, const void*, int64_t, int64_t,
                   const void*, int64_t, int64_t, int, int, void*, int, int*, int*) { return 1; }
int svd_batch_run(const void*, int, int, int, int, int, cudaStream_t) { return set_error(SDNQ_EUNSUPPORTED, "the batched tensor-core dequant kernel is not emulated"); }
}  // namespace sdnq

extern "C" {
int sdnq_b200_abi_version(void) { return SDNQ_B200_ABI_VERSION; }
const char* sdnq_b200_last_error(void) { return sdnq::g_error; }
int sdnq_b200_check_device(int) { return SDNQ_OK; }
int64_t sdnq_b200_launch_count(int reset) {
    const int64_t v = sdnq::g_launches;
    if (reset) sdnq::g_launches = 0;
    return v;
}
int64_t sdnq_b200_stream_capture_id(void*) { return 0; }      // the emulator has no streams: never capturing
// the two CUDA runtime calls the launch code makes besides the launch itself
const char* cudaGetErrorString(cudaError_t) { return "cuda runtime is not available in the host emulator"; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
}

extern "C" {

// storage bytes -> unsigned codes, octet by octet, through load_octet_bytes + decode_octet exactly as the kernels do
int emu_decode(int bits, int word_bytes, const uint8_t* packed, int64_t octets, uint32_t* codes) {
    if (bits < 1 || bits > 8) return -1;
    for (int64_t o = 0; o < octets; ++o) {
        SDNQ_DISPATCH_BITS(bits, {
            uint32_t w[OctetWords<BITS>::N];
            load_octet_bytes<BITS>(packed, o, word_bytes, w);
            uint32_t v[8];
            decode_octet<BITS>(w, v);
            for (int i = 0; i < 8; ++i) codes[8 * o + i] = v[i];
        });
    }
    return 0;
}

// storage bytes -> real values before the scale (codes_to_values: integer offset / minifloat / fp8 decode)
int emu_values(const sdnq_weight_format* fmt, const uint8_t* packed, int64_t octets, float* out) {
    WFormat f;
    int rc = make_wformat(fmt, &f);
    if (rc != SDNQ_OK) return rc;
    for (int64_t o = 0; o < octets; ++o) {
        SDNQ_DISPATCH_BITS(f.bits, {
            float q[8];
            uint32_t c[8];
            octet_values<BITS>(packed, o, f, q, c);
            for (int i = 0; i < 8; ++i) out[8 * o + i] = q[i];
        });
    }
    return 0;
}

// the byte-permute fast path of the flat dequant kernel / GEMV for integer formats
int emu_octet_to_floats(const sdnq_weight_format* fmt, const uint8_t* packed, int64_t octets, float* out) {
    WFormat f;
    int rc = make_wformat(fmt, &f);
    if (rc != SDNQ_OK || f.kind != SDNQ_W_INT) return -1;
    const bool twos = f.bits == 8 && !f.is_unsigned;
    const uint32_t flip = twos ? 0x80808080u : 0u;
    const float bias = twos ? 8388608.0f + 128.0f : 8388608.0f - static_cast<float>(f.int_offset);
    for (int64_t o = 0; o < octets; ++o) {
        SDNQ_DISPATCH_BITS(f.bits, {
            uint32_t w[OctetWords<BITS>::N];
            load_octet_bytes<BITS>(packed, o, f.word_bytes, w);
            float q[8];
            octet_to_floats<BITS>(w, flip, bias, q);
            for (int i = 0; i < 8; ++i) out[8 * o + i] = q[i];
        });
    }
    return 0;
}

void emu_f32_to_e4m3(const float* in, int64_t n, uint8_t* out) { for (int64_t i = 0; i < n; ++i) out[i] = f32_to_e4m3(in[i]); }
void emu_e4m3_to_f32(const uint8_t* in, int64_t n, float* out) { for (int64_t i = 0; i < n; ++i) out[i] = e4m3_to_f32(in[i]); }
void emu_e5m2_to_f32(const uint8_t* in, int64_t n, float* out) { for (int64_t i = 0; i < n; ++i) out[i] = e5m2_to_f32(in[i]); }
void emu_round_bf16(const float* in, int64_t n, float* out) { for (int64_t i = 0; i < n; ++i) out[i] = ElemTraits<__nv_bfloat16>::round(in[i]); }
void emu_round_f16(const float* in, int64_t n, float* out) { for (int64_t i = 0; i < n; ++i) out[i] = ElemTraits<__half>::round(in[i]); }

// where the two 4-element halves of lane `lane` land after the in-warp H4-family transform, and the per-register sign rule
int emu_hadamard_dest(int G, int lane, int half) { return hadamard_dest_dyn(G, lane, half); }
uint32_t emu_hadamard_sign(int G, int lane, int j) {
    switch (G) {
        case 4: return hadamard_sign_mask<4>(lane, j);
        case 16: return hadamard_sign_mask<16>(lane, j);
        case 64: return hadamard_sign_mask<64>(lane, j);
        case 256: return hadamard_sign_mask<256>(lane, j);
        default: return 0;
    }
}
}

// ---------------------------------------------------------------- warp-level code on 32 lock-stepped host threads
#include "../../sdnq_b200/csrc/hadamard_tc.cuh"

namespace {
template <typename T>
void rotate_tc(int G, const uint16_t* in, uint16_t* out, float* lane_max, int64_t chunks) {
    sdnq_emu::run_warp([&](int lane) {
        hadtc::Rotation<T> rot;
        rot.init(G, lane);
        const float factor = hadamard_factor<T>(G);
        for (int64_t c = 0; c < chunks; ++c) {
            // the kernels' two coalesced 8-byte accesses: elements [4l, 4l+4) and [128+4l, 128+4l+4) of the 256-chunk
            const uint16_t* p = in + 256 * c;
            uint4 raw;
            std::memcpy(&raw.x, p + 4 * lane, 8);
            std::memcpy(&raw.z, p + 128 + 4 * lane, 8);
            const float m = rot.apply(raw, factor);
            std::memcpy(out + 256 * c + 4 * lane, &raw.x, 8);
            std::memcpy(out + 256 * c + 128 + 4 * lane, &raw.z, 8);
            lane_max[32 * c + lane] = m;
        }
    });
}

template <typename T>
void rotate_butterfly(int G, const float* in, float* out, int64_t chunks) {
    sdnq_emu::run_warp([&](int lane) {
        const float factor = hadamard_factor<T>(G);
        for (int64_t c = 0; c < chunks; ++c) {
            float v[8];
            for (int i = 0; i < 8; ++i) v[i] = in[256 * c + 8 * lane + i];
            hadamard_warp_dyn(G, v, factor);
            for (int half = 0; half < 2; ++half) {
                const int dst = hadamard_dest_dyn(G, lane, half);
                for (int i = 0; i < 4; ++i) out[256 * c + dst + i] = v[4 * half + i];
            }
        }
    });
}
}  // namespace

extern "C" {
// Rotation<T>::apply (tensor-core path of K2 / rotated K3) over `chunks` 256-element chunks of 16-bit values
void emu_rotate_tc(int f16, int G, const uint16_t* in, uint16_t* out, float* lane_max, int64_t chunks) {
    if (f16) rotate_tc<__half>(G, in, out, lane_max, chunks);
    else rotate_tc<__nv_bfloat16>(G, in, out, lane_max, chunks);
}
// hadamard_warp<G> + hadamard_dest<G> (shuffle-butterfly path), values scaled by the factor in T, not yet rounded to T
void emu_rotate_butterfly(int f16, int G, const float* in, float* out, int64_t chunks) {
    if (f16) rotate_butterfly<__half>(G, in, out, chunks);
    else rotate_butterfly<__nv_bfloat16>(G, in, out, chunks);
}
}

// ---------------------------------------------------------------- a whole kernel: K5p (gemv_packed_kernel.cuh) on emulated CTAs
#include "../../sdnq_b200/csrc/gemv_packed_kernel.cuh"

namespace {
template <typename T, int BITS>
void run_gemv_packed(const gemvp::Args& a, int grid) {
    const int mb = (a.M + 7) / 8;
    static float s_red[(gemvp::kWarps - 1) * 4 * 4 * 32];        // the kernel's __shared__ array (CTAs run one at a time)
    sdnq_emu::run_grid(grid, gemvp::kThreads, [&] {
        if (mb <= 1) gemvp::body<T, BITS, 1>(a, s_red);
        else if (mb == 2) gemvp::body<T, BITS, 2>(a, s_red);
        else gemvp::body<T, BITS, 4>(a, s_red);
    });
}
}  // namespace

extern "C" int emu_gemv_packed(const void* x, int x_dtype, int64_t ldx, const void* weight, const sdnq_weight_format* fmt,
                               const float* scale, const float* zero_point, int64_t group_size, const void* bias, int bias_dtype,
                               int64_t bias_ld, void* out, int64_t M, int64_t N, int64_t K, int grid) {
    WFormat f;
    int rc = make_wformat(fmt, &f);
    if (rc != SDNQ_OK) return rc;
    const int64_t group = (group_size <= 0 || group_size >= K) ? K : group_size;
    if (group % 8 != 0 || K % group != 0 || K % 16 != 0 || M < 1 || M > 32) return -1;
    gemvp::Args a{x, ldx, reinterpret_cast<const uint8_t*>(weight), scale, zero_point, bias, bias_dtype, bias_ld, out,
                  int(M), int(N), int(K), int(group), int(K / group), f};
    if (x_dtype == SDNQ_BF16) { SDNQ_DISPATCH_BITS(f.bits, run_gemv_packed<__nv_bfloat16, BITS>(a, grid)); }
    else { SDNQ_DISPATCH_BITS(f.bits, run_gemv_packed<__half, BITS>(a, grid)); }
    return 0;
}
