// K2: activation pre-pass -- Hadamard rotation + per-row dynamic quantisation.
//
// Reference behaviour restated here:
//   rotate_hadamard                      quant_utils.py:193-209   (x.unflatten(-1,(-1,g)) @ H in x.dtype)
//   quantize_int_mm / uint_mm / fp_mm    quant_utils.py:264-299   (true f32 division, round-half-even)
//   quantize_*_mm_input                  layers/linear/linear_int8.py:14-22, linear_uint8.py:14-23, linear_fp8.py:14-22
//   zero-point row sums                  layers/linear/linear_int8.py:65-69
//
// Layout: a row is cut into 256-element chunks; a warp owns whole chunks (lane l holds elements [8l, 8l+8) of
// the chunk in registers), so every global access is 16 B per lane / 512 B contiguous per warp and the whole
// row stays in registers between the amax pass and the quantise pass: x is read from HBM exactly once
// (2 B/elem in) and xq written once (1 B/elem out).  WPR warps cooperate on one row (cross-warp amax through
// shared memory), 8/WPR rows per CTA.
#include "common.cuh"

namespace sdnq {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = 8;

struct ActArgs {
    const void* x;
    int64_t M, K, ldx;
    int hadamard;
    int mode;            // SDNQ_I8 / SDNQ_U8 / SDNQ_F8E4M3
    uint8_t* xq;
    float* sx;
    float* zx;
    int32_t* rowsum;
    void* x_rot;
};

template <typename T, int WPR, int MAXC>
__global__ void __launch_bounds__(kThreads) act_quant_kernel(const ActArgs a) {
    constexpr int RPC = kWarps / WPR;                 // rows per CTA
    __shared__ float s_a[RPC][WPR];
    __shared__ float s_b[RPC][WPR];
    __shared__ int s_sum[RPC][WPR];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r_in = warp / WPR, w_in = warp % WPR;
    const int64_t row = int64_t(blockIdx.x) * RPC + r_in;
    const bool row_ok = row < a.M;
    const T* xrow = reinterpret_cast<const T*>(a.x) + row * a.ldx;

    float v[MAXC][8];
    float amax = 0.f, vmax = -INFINITY, vmin = INFINITY;
    const float hfac = a.hadamard ? hadamard_factor<T>(a.hadamard) : 1.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int64_t k0 = (int64_t(c) * WPR + w_in) * 256;          // chunk start (warp-uniform)
        const int64_t k = k0 + lane * 8;
        const bool ok = row_ok && k < a.K;
        if (ok) {
            load8<T>(xrow + k, v[c]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[c][i] = 0.f;
        }
        if (a.hadamard && k0 < a.K) {
            hadamard_warp_dyn(a.hadamard, v[c]);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[c][i] = ElemTraits<T>::round(v[c][i] * hfac);   // result of the matmul is in x.dtype
        }
        if (ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                amax = fmaxf(amax, fabsf(v[c][i]));
                vmax = fmaxf(vmax, v[c][i]);
                vmin = fminf(vmin, v[c][i]);
            }
        }
    }
    // ---- row statistics
    float scale, zero = 0.f;
    if (a.mode == SDNQ_U8) {
        vmax = warp_max(vmax);
        vmin = warp_min(vmin);
        if (WPR > 1) {
            if (lane == 0) { s_a[r_in][w_in] = vmax; s_b[r_in][w_in] = vmin; }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < WPR; ++i) { vmax = fmaxf(vmax, s_a[r_in][i]); vmin = fminf(vmin, s_b[r_in][i]); }
        }
        scale = __fdiv_rn(__fsub_rn(vmax, vmin), 255.f);                 // get_scale_asymmetric(.., "int8")
        zero = __fsub_rn(vmin, __fmul_rn(scale, -128.f));
    } else {
        amax = warp_max(amax);
        if (WPR > 1) {
            if (lane == 0) s_a[r_in][w_in] = amax;
            __syncthreads();
#pragma unroll
            for (int i = 0; i < WPR; ++i) amax = fmaxf(amax, s_a[r_in][i]);
        }
        scale = __fdiv_rn(amax, a.mode == SDNQ_F8E4M3 ? 448.f : 127.f);  // get_scale_symmetric
    }
    // ---- quantise from registers
    int local_sum = 0;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int64_t k = (int64_t(c) * WPR + w_in) * 256 + lane * 8;
        if (!(row_ok && k < a.K)) continue;
        if (a.x_rot != nullptr) store8<T>(reinterpret_cast<T*>(a.x_rot) + row * a.K + k, v[c]);
        uint2 r;
        uint8_t* b = reinterpret_cast<uint8_t*>(&r);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float q = v[c][i];
            if (a.mode == SDNQ_U8) q = __fsub_rn(q, zero);
            q = __fdiv_rn(q, scale);
            if (a.mode == SDNQ_F8E4M3) {
                if (q != q) q = 0.f;                                     // nan_to_num
                b[i] = f32_to_e4m3(fminf(fmaxf(q, -448.f), 448.f));
            } else {
                q = rintf(q);
                const int ci = (q != q) ? 0 : static_cast<int>(fminf(fmaxf(q, -128.f), 127.f));
                b[i] = static_cast<uint8_t>(static_cast<int8_t>(ci));
                local_sum += ci;
            }
        }
        *reinterpret_cast<uint2*>(a.xq + row * a.K + k) = r;
    }
    if (a.rowsum != nullptr) {
        local_sum = warp_sum(local_sum);
        if (WPR > 1) {
            if (lane == 0) s_sum[r_in][w_in] = local_sum;
            __syncthreads();
            local_sum = 0;
#pragma unroll
            for (int i = 0; i < WPR; ++i) local_sum += s_sum[r_in][i];
        }
        if (row_ok && w_in == 0 && lane == 0) a.rowsum[row] = local_sum;
    }
    if (row_ok && w_in == 0 && lane == 0) {
        a.sx[row] = scale;
        if (a.zx != nullptr) a.zx[row] = zero;
    }
}

template <typename T, int WPR, int MAXC>
int launch(const ActArgs& a, cudaStream_t st) {
    constexpr int RPC = kWarps / WPR;
    const unsigned blocks = static_cast<unsigned>((a.M + RPC - 1) / RPC);
    act_quant_kernel<T, WPR, MAXC><<<blocks, kThreads, 0, st>>>(a);
    return check_launch("act_quant_kernel");
}

template <typename T>
int dispatch(const ActArgs& a, cudaStream_t st) {
    const int64_t chunks = (a.K + 255) / 256;
    if (chunks <= 1) return launch<T, 1, 1>(a, st);
    if (chunks <= 2) return launch<T, 1, 2>(a, st);
    if (chunks <= 4) return launch<T, 1, 4>(a, st);
    if (chunks <= 8) return launch<T, 2, 4>(a, st);
    if (chunks <= 16) return launch<T, 4, 4>(a, st);
    if (chunks <= 32) return launch<T, 8, 4>(a, st);
    if (chunks <= 64) return launch<T, 8, 8>(a, st);
    return set_error(SDNQ_EUNSUPPORTED, "act_quant: K=%lld exceeds 16384", (long long)a.K);
}

}  // namespace

int act_quant_impl(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx, int hadamard_group, int mm_dtype, void* xq,
                   float* sx, float* zx, int32_t* rowsum, void* x_rot, cudaStream_t st) {
    SDNQ_REQUIRE(x && xq && sx, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(M >= 0 && K > 0 && ldx >= K, SDNQ_EINVAL, "bad shape M=%lld K=%lld ldx=%lld", (long long)M, (long long)K, (long long)ldx);
    SDNQ_REQUIRE(K % 8 == 0 && ldx % 8 == 0, SDNQ_EUNSUPPORTED, "K and ldx must be multiples of 8 (K=%lld ldx=%lld)", (long long)K, (long long)ldx);
    SDNQ_REQUIRE(mm_dtype == SDNQ_I8 || mm_dtype == SDNQ_U8 || mm_dtype == SDNQ_F8E4M3, SDNQ_EINVAL, "bad matmul dtype %d", mm_dtype);
    SDNQ_REQUIRE(mm_dtype != SDNQ_U8 || zx != nullptr, SDNQ_EINVAL, "uint8 activations need a zx output");
    SDNQ_REQUIRE(hadamard_group == 0 || (hadamard_group >= 4 && hadamard_group <= 256 && (hadamard_group & (hadamard_group - 1)) == 0),
                 SDNQ_EUNSUPPORTED, "hadamard group %d: only powers of two in [4,256] are implemented", hadamard_group);
    if (hadamard_group) SDNQ_REQUIRE(K % hadamard_group == 0, SDNQ_EINVAL, "hadamard group %d does not divide K=%lld", hadamard_group, (long long)K);
    SDNQ_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(xq) & 7) == 0, SDNQ_EINVAL, "x must be 16-byte and xq 8-byte aligned");
    if (M == 0) return SDNQ_OK;
    ActArgs a{x, M, K, ldx, hadamard_group, mm_dtype, reinterpret_cast<uint8_t*>(xq), sx, mm_dtype == SDNQ_U8 ? zx : nullptr, rowsum, x_rot};
    switch (x_dtype) {
        case SDNQ_BF16: return dispatch<__nv_bfloat16>(a, st);
        case SDNQ_F16: return dispatch<__half>(a, st);
        case SDNQ_F32: return dispatch<float>(a, st);
        default: return set_error(SDNQ_EINVAL, "bad activation dtype %d", x_dtype);
    }
}

}  // namespace sdnq

extern "C" int sdnq_b200_act_quant(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx, int hadamard_group, int mm_dtype,
                                   void* xq, float* sx, float* zx, int32_t* rowsum, void* x_rot, void* stream) {
    return sdnq::act_quant_impl(x, x_dtype, M, K, ldx, hadamard_group, mm_dtype, xq, sx, zx, rowsum, x_rot,
                                reinterpret_cast<cudaStream_t>(stream));
}
