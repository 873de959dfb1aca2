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
#include "act_quant.cuh"

namespace sdnq {
namespace {

using namespace actq;

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

template <typename T, int WPR, int MAXC, int MODE>
__global__ void __launch_bounds__(kThreads, (sizeof(T) == 2 && MAXC <= 4) ? 6 : 3) act_quant_kernel(const ActArgs a) {
    constexpr int RPC = kWarps / WPR;                 // rows per CTA
    __shared__ float s_a[RPC][WPR];
    __shared__ float s_b[RPC][WPR];
    __shared__ int s_sum[RPC][WPR];
    pdl_launch_dependents();      // the GEMM behind us may start its prologue / weight prefetch now
    pdl_wait();                   // x (and the workspace we overwrite) belong to the stream predecessor
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r_in = warp / WPR, w_in = warp % WPR;
    const int64_t row = int64_t(blockIdx.x) * RPC + r_in;
    const bool row_ok = row < a.M;
    const T* xrow = reinterpret_cast<const T*>(a.x) + row * a.ldx;
    const int K = static_cast<int>(a.K);

    Held<T> held[MAXC];
    // all loads of the row first (memory-level parallelism), statistics afterwards
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int k = (c * WPR + w_in) * 256 + lane * 8;
        if (row_ok && k < K) held[c].load(xrow + k);
        else held[c].zero();
    }
    float amax = 0.f, vmax = -INFINITY, vmin = INFINITY;
    const float hfac = a.hadamard ? hadamard_factor<T>(a.hadamard) : 1.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int k0 = (c * WPR + w_in) * 256;          // chunk start (warp-uniform)
        const bool ok = row_ok && k0 + lane * 8 < K;
        float v[8];
        held[c].get(v);
        if (a.hadamard && k0 < K) {
            hadamard_warp_dyn(a.hadamard, v, hfac);            // put() rounds to x.dtype: the reference's matmul returns x.dtype
            held[c].put(v);
            held[c].get(v);
        }
        if (ok) {
            if constexpr (MODE == SDNQ_U8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { vmax = fmaxf(vmax, v[i]); vmin = fminf(vmin, v[i]); }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) amax = fmaxf(amax, fabsf(v[i]));
            }
        }
    }
    // ---- row statistics
    float scale, zero = 0.f;
    if constexpr (MODE == SDNQ_U8) {
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
        scale = __fdiv_rn(amax, MODE == SDNQ_F8E4M3 ? 448.f : 127.f);    // get_scale_symmetric
    }
    const RowDivider divider(scale);
    const bool safe = divider.safe();                                    // uniform across the row (and the warp)
    // ---- quantise from registers
    int local_sum = 0;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int k = (c * WPR + w_in) * 256 + lane * 8;
        if (!(row_ok && k < K)) continue;
        float v[8];
        held[c].get(v);
        const bool want_sum = a.rowsum != nullptr;
        const uint2 r = safe ? quantise8<MODE, true>(v, divider, zero, want_sum, local_sum) : quantise8<MODE, false>(v, divider, zero, want_sum, local_sum);
        // after a power-of-4 Hadamard the lane's two 4-element halves belong elsewhere in the chunk (see hadamard_dest)
        const int64_t chunk0 = row * a.K + (k - lane * 8);
        const int d0 = hadamard_dest_dyn(a.hadamard, lane, 0), d1 = hadamard_dest_dyn(a.hadamard, lane, 1);
        *reinterpret_cast<uint32_t*>(a.xq + chunk0 + d0) = r.x;
        *reinterpret_cast<uint32_t*>(a.xq + chunk0 + d1) = r.y;
        if (a.x_rot != nullptr) {
            T* xr = reinterpret_cast<T*>(a.x_rot) + chunk0;
            store4<T>(xr + d0, v[0], v[1], v[2], v[3]);
            store4<T>(xr + d1, v[4], v[5], v[6], v[7]);
        }
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
    cudaError_t e;
    if (a.mode == SDNQ_I8) e = launch_pdl(act_quant_kernel<T, WPR, MAXC, SDNQ_I8>, dim3(blocks), dim3(kThreads), 0, st, a);
    else if (a.mode == SDNQ_U8) e = launch_pdl(act_quant_kernel<T, WPR, MAXC, SDNQ_U8>, dim3(blocks), dim3(kThreads), 0, st, a);
    else e = launch_pdl(act_quant_kernel<T, WPR, MAXC, SDNQ_F8E4M3>, dim3(blocks), dim3(kThreads), 0, st, a);
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of act_quant_kernel failed: %s", cudaGetErrorString(e));
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
