// K2 kernel template (shared by act_quant.cu = Linear rows and act_quant_conv.cu = im2col rows of a convolution; two
// translation units so the two instantiation sets compile in parallel and the gather code never touches the Linear kernels).
#pragma once
#include "act_quant.cuh"
#include "hadamard_tc.cuh"

#include <cmath>
#include <cstdlib>
#include <type_traits>

namespace sdnq {
namespace {

using namespace actq;

constexpr int kThreads = 256;
constexpr int kWarps = 8;

// im2col view of a convolution input (conv2d; conv1d as H = 1): row m = (b, oh, ow), column k = (c, i, j) with j fastest --
// the order of F.unfold(...).transpose(1, 2) in process_conv_input (layers/conv/forward.py:30-76).  Never materialised:
// the quantiser gathers its row straight from the NCHW (or any strided) input, padding reads as zero.
struct ConvView {
    int on;
    int C, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Wout, HWout;
    int64_t sB, sC, sH, sW;      // element strides of the input tensor
};

struct ActArgs {
    const void* x;
    int64_t M, K, ldx;
    int hadamard;
    float hfac;          // 1/sqrt(hadamard) rounded to the activation dtype (H.div_(n**0.5) in x.dtype, quant_utils.py:151,163)
    int mode;            // SDNQ_I8 / SDNQ_U8 / SDNQ_F8E4M3
    uint8_t* xq;
    float* sx;
    float* zx;
    int32_t* rowsum;
    void* x_rot;
    ConvView conv;
};

// n consecutive im2col columns starting at k0 of the row whose window origin is (ih0, iw0) in image `img`
template <typename T, int N>
__device__ __noinline__ void conv_gather(const ConvView& cv, const T* __restrict__ img, int ih0, int iw0, int k0, int K, T (&out)[N]) {
    const int taps = cv.kh * cv.kw;
    int ch = k0 / taps;
    const int t = k0 - ch * taps;
    int i = t / cv.kw, j = t - i * cv.kw;
#pragma unroll
    for (int e = 0; e < N; ++e) {
        const int ih = ih0 + i * cv.dh, iw = iw0 + j * cv.dw;
        T v = T(0.0f);
        if (k0 + e < K && ih >= 0 && ih < cv.H && iw >= 0 && iw < cv.W) v = img[ch * cv.sC + ih * cv.sH + iw * cv.sW];
        out[e] = v;
        if (++j == cv.kw) { j = 0; if (++i == cv.kh) { i = 0; ++ch; } }
    }
}
template <typename T>
__device__ __forceinline__ uint32_t pack16(T a, T b) {
    return uint32_t(*reinterpret_cast<const uint16_t*>(&a)) | (uint32_t(*reinterpret_cast<const uint16_t*>(&b)) << 16);
}

// kTC: the rotation runs on the tensor cores (hadamard_tc.cuh; 16-bit activations).  The lane then owns elements
// [4l, 4l+4) and [128+4l, 128+4l+4) of a chunk (two coalesced 8-byte accesses) instead of [8l, 8l+8).
template <typename T, int WPR, int MAXC, int MODE, bool kTC, bool kConv>
__global__ void __launch_bounds__(kThreads, (sizeof(T) == 2 && MAXC <= 4) ? (kTC ? 4 : 5) : (kTC ? 2 : 3)) act_quant_kernel(const ActArgs a) {
    constexpr int RPC = kWarps / WPR;                 // rows per CTA and pass
    // the cross-warp exchange buffers alternate between passes of the row loop: one barrier per pass is enough
    __shared__ float s_a2[2][RPC][WPR];
    __shared__ float s_b2[2][RPC][WPR];
    __shared__ int s_sum2[2][RPC][WPR];
    pdl_launch_dependents();      // the GEMM behind us may start its prologue / weight prefetch now
    pdl_wait();                   // x (and the workspace we overwrite) belong to the stream predecessor
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r_in = warp / WPR, w_in = warp % WPR;
    const float hfac = a.hfac;
    [[maybe_unused]] hadtc::Rotation<std::conditional_t<kTC, T, __nv_bfloat16>> rot;
    if constexpr (kTC) rot.init(a.hadamard, lane);
    // Row loop: the grid is a few CTAs per SM and every CTA walks row blocks blockIdx.x, blockIdx.x + gridDim.x, ... -- the per-CTA
    // set-up (rotation constants, dependency wait, argument loads) is paid once, not once per pair of rows.
    const int64_t nblk = (a.M + RPC - 1) / RPC;
    // Long Linear rows of 16-bit activations: the NEXT row block's chunks are loaded before this one is reduced, quantised and stored
    // (register double buffer), so the SM keeps loads in flight through the statistics barrier and the store phase -- without it
    // every CTA alternates between a load burst and a phase with nothing outstanding, and the kernel sits at ~0.55 of the copy rate.
    // Measured on B200 (tools/actq_grid.py): rows of 8k+ elements (one row per CTA, 6 chunks per warp) gain 15 % (3.05 -> 3.56 TB/s at
    // 16384 x 12288 with Hadamard-256); the 4-chunk variants lose more to the extra registers (one resident CTA less) than they gain.
    constexpr bool kPrefetch = !kConv && sizeof(T) == 2 && MAXC > 4;
    constexpr int kStep = WPR * 256;                                  // column distance between this warp's consecutive chunks
    const int kb = w_in * 256 + lane * (kTC ? 4 : 8);                 // this lane's first column
    auto load_linear = [&](int64_t blk_i, Held<T> (&h)[MAXC]) {
        const int64_t row_i = blk_i * RPC + r_in;
        const int lim_i = (blk_i < nblk && row_i < a.M) ? static_cast<int>(a.K) - kb : 0;
        const T* xp_i = reinterpret_cast<const T*>(a.x) + row_i * a.ldx + kb;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            if constexpr (kTC) {
                uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
                if (c * kStep < lim_i) lo = *reinterpret_cast<const uint2*>(xp_i + c * kStep);
                if (c * kStep + 128 < lim_i) hi = *reinterpret_cast<const uint2*>(xp_i + c * kStep + 128);
                h[c].raw = make_uint4(lo.x, lo.y, hi.x, hi.y);
            } else {
                if (c * kStep < lim_i) h[c].load(xp_i + c * kStep);
                else h[c].zero();
            }
        }
    };
    Held<T> held[MAXC];
    [[maybe_unused]] Held<T> ahead[MAXC];
    if constexpr (kPrefetch) load_linear(blockIdx.x, held);
    int pass = 0;
    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x, pass ^= 1) {
    if constexpr (kPrefetch) load_linear(blk + gridDim.x, ahead);
    float (&s_a)[RPC][WPR] = s_a2[pass];
    float (&s_b)[RPC][WPR] = s_b2[pass];
    int (&s_sum)[RPC][WPR] = s_sum2[pass];
    const int64_t row = blk * RPC + r_in;
    const bool row_ok = row < a.M;
    const T* xrow = reinterpret_cast<const T*>(a.x) + (kConv ? 0 : row * a.ldx);
    const int K = static_cast<int>(a.K);
    int ih0 = 0, iw0 = 0;
    if (kConv && row_ok) {
        const int b = static_cast<int>(row / a.conv.HWout);
        const int rem = static_cast<int>(row - int64_t(b) * a.conv.HWout);
        const int oh = rem / a.conv.Wout, ow = rem - oh * a.conv.Wout;
        ih0 = oh * a.conv.sh - a.conv.ph;
        iw0 = ow * a.conv.sw - a.conv.pw;
        xrow += int64_t(b) * a.conv.sB;
    }

    // Per-thread bases; inside the unrolled loops every address is base + a compile-time offset and every validity test is
    // one compare against a constant (hoisted by hand: under the register cap the compiler re-derived them per chunk).
    const int lim = row_ok ? K - kb : 0;                              // chunk c holds data for this lane iff c * kStep < lim
    const int wlim = row_ok ? K - w_in * 256 : 0;                     // ... for this warp (uniform)
    const T* xp = xrow + (kConv ? 0 : kb);
    uint8_t* qp = a.xq + row * a.K + kb;
    T* rp = a.x_rot != nullptr ? reinterpret_cast<T*>(a.x_rot) + row * a.K + kb : nullptr;

    // all loads of the row first (memory-level parallelism), statistics afterwards
    if constexpr (!kPrefetch)
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        if constexpr (kTC) {
            uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
            if constexpr (kConv) {
                T g[4];
                if (c * kStep < lim) { conv_gather<T, 4>(a.conv, xrow, ih0, iw0, kb + c * kStep, K, g); lo = make_uint2(pack16(g[0], g[1]), pack16(g[2], g[3])); }
                if (c * kStep + 128 < lim) { conv_gather<T, 4>(a.conv, xrow, ih0, iw0, kb + c * kStep + 128, K, g); hi = make_uint2(pack16(g[0], g[1]), pack16(g[2], g[3])); }
            } else {
                if (c * kStep < lim) lo = *reinterpret_cast<const uint2*>(xp + c * kStep);
                if (c * kStep + 128 < lim) hi = *reinterpret_cast<const uint2*>(xp + c * kStep + 128);
            }
            held[c].raw = make_uint4(lo.x, lo.y, hi.x, hi.y);
        } else {
            if (!(c * kStep < lim)) held[c].zero();
            else if constexpr (kConv) {
                T g[8];
                conv_gather<T, 8>(a.conv, xrow, ih0, iw0, kb + c * kStep, K, g);
                if constexpr (sizeof(T) == 2) held[c].raw = make_uint4(pack16(g[0], g[1]), pack16(g[2], g[3]), pack16(g[4], g[5]), pack16(g[6], g[7]));
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) held[c].val[i] = g[i];
                }
            } else held[c].load(xp + c * kStep);
        }
    }
    float amax = 0.f, vmax = -INFINITY, vmin = INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        float v[8];
        if constexpr (kTC) {
            float m = 0.f;
            if (c * kStep < wlim) m = rot.apply(held[c].raw, hfac);    // warp-uniform; rounds to x.dtype: the reference's matmul returns x.dtype
            if constexpr (MODE == SDNQ_U8) {
                held[c].get(v);
                const bool ok_lo = c * kStep < lim, ok_hi = c * kStep + 128 < lim;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < 4 ? ok_lo : ok_hi) { vmax = fmaxf(vmax, v[i]); vmin = fminf(vmin, v[i]); }
                }
            } else {     // halves past the end of the row hold zeros (groups never straddle the end): no effect on an absolute maximum;
                         // the maximum of the unrounded values is rounded once below (monotone rounding commutes with max)
                amax = fmaxf(amax, m);
            }
            continue;
        }
        const bool ok = c * kStep < lim;
        held[c].get(v);
        if (a.hadamard && c * kStep < wlim) {
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
        if constexpr (kTC) amax = ElemTraits<T>::round(amax);            // the rotated values are stored rounded to x.dtype
        scale = __fdiv_rn(amax, MODE == SDNQ_F8E4M3 ? 448.f : 127.f);    // get_scale_symmetric
    }
    const RowDivider divider(scale);
    const bool safe = divider.safe();                                    // uniform across the row (and the warp)
    // ---- quantise from registers
    int local_sum = 0;
    [[maybe_unused]] const int d0 = kTC ? 0 : hadamard_dest_dyn(a.hadamard, lane, 0) - lane * 8;
    [[maybe_unused]] const int d1 = kTC ? 4 : hadamard_dest_dyn(a.hadamard, lane, 1) - lane * 8;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        if constexpr (kTC) {
            if (!(c * kStep < lim)) continue;
            const bool ok_hi = c * kStep + 128 < lim;
            float v[8];
            held[c].get(v);
            int sum8 = 0;
            const bool want_sum = a.rowsum != nullptr;
            const uint2 r = safe ? quantise8<MODE, true>(v, divider, zero, want_sum, sum8) : quantise8<MODE, false>(v, divider, zero, want_sum, sum8);
            if (want_sum) {
                if (ok_hi) local_sum += sum8;
                else {      // only the low half exists: recount its four codes
                    const int8_t* cb = reinterpret_cast<const int8_t*>(&r.x);
                    local_sum += cb[0] + cb[1] + cb[2] + cb[3];
                }
            }
            *reinterpret_cast<uint32_t*>(qp + c * kStep) = r.x;
            if (ok_hi) *reinterpret_cast<uint32_t*>(qp + c * kStep + 128) = r.y;
            if (rp != nullptr) {
                *reinterpret_cast<uint2*>(rp + c * kStep) = make_uint2(held[c].raw.x, held[c].raw.y);
                if (ok_hi) *reinterpret_cast<uint2*>(rp + c * kStep + 128) = make_uint2(held[c].raw.z, held[c].raw.w);
            }
            continue;
        }
        if (!(c * kStep < lim)) continue;
        float v[8];
        held[c].get(v);
        const bool want_sum = a.rowsum != nullptr;
        const uint2 r = safe ? quantise8<MODE, true>(v, divider, zero, want_sum, local_sum) : quantise8<MODE, false>(v, divider, zero, want_sum, local_sum);
        // after a power-of-4 Hadamard the lane's two 4-element halves belong elsewhere in the chunk (see hadamard_dest)
        // d0 / d1: where the lane's two 4-element halves belong, relative to its own column (0 / 4 unless a power-of-4 Hadamard
        // left them permuted inside the chunk, see hadamard_dest)
        *reinterpret_cast<uint32_t*>(qp + c * kStep + d0) = r.x;
        *reinterpret_cast<uint32_t*>(qp + c * kStep + d1) = r.y;
        if (rp != nullptr) {
            store4<T>(rp + c * kStep + d0, v[0], v[1], v[2], v[3]);
            store4<T>(rp + c * kStep + d1, v[4], v[5], v[6], v[7]);
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
    if constexpr (kPrefetch) {
#pragma unroll
        for (int c = 0; c < MAXC; ++c) held[c] = ahead[c];
    }
    }   // row loop
}

// Rows longer than the register-resident kernel holds (K > 16384): one CTA per row, two passes over the row (statistics, then
// quantise; the second read comes from L2).  Same arithmetic, same layout (lane l owns elements [8l, 8l+8) of a 256-chunk, warp w
// takes chunks w, w+8, ...); the butterfly rotation is applied in both passes when asked for.
template <typename T, int MODE, bool kConv>
__global__ void __launch_bounds__(kThreads) act_quant_long_kernel(const ActArgs a) {
    __shared__ float s_a[kWarps];
    __shared__ float s_b[kWarps];
    __shared__ int s_sum[kWarps];
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = blockIdx.x;
    const int K = static_cast<int>(a.K);
    const T* xrow = reinterpret_cast<const T*>(a.x) + (kConv ? 0 : row * a.ldx);
    int ih0 = 0, iw0 = 0;
    if constexpr (kConv) {
        const int b = static_cast<int>(row / a.conv.HWout);
        const int rem = static_cast<int>(row - int64_t(b) * a.conv.HWout);
        const int oh = rem / a.conv.Wout, ow = rem - oh * a.conv.Wout;
        ih0 = oh * a.conv.sh - a.conv.ph;
        iw0 = ow * a.conv.sw - a.conv.pw;
        xrow += int64_t(b) * a.conv.sB;
    }
    const int chunks = (K + 255) / 256;
    auto fetch = [&](int c, float (&v)[8]) {          // the lane's 8 values of chunk c (rotated if asked), zeros past the end
        const int k = c * 256 + lane * 8;
        Held<T> h;
        if (k >= K) h.zero();
        else if constexpr (kConv) {
            T g[8];
            conv_gather<T, 8>(a.conv, xrow, ih0, iw0, k, K, g);
            if constexpr (sizeof(T) == 2) h.raw = make_uint4(pack16(g[0], g[1]), pack16(g[2], g[3]), pack16(g[4], g[5]), pack16(g[6], g[7]));
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) h.val[i] = g[i];
            }
        } else h.load(xrow + k);
        h.get(v);
        if (a.hadamard) {
            hadamard_warp_dyn(a.hadamard, v, a.hfac);
            h.put(v);                                   // rounds to x.dtype
            h.get(v);
        }
    };
    float amax = 0.f, vmax = -INFINITY, vmin = INFINITY;
    for (int c = warp; c < chunks; c += kWarps) {
        float v[8];
        fetch(c, v);
        if (c * 256 + lane * 8 < K) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if constexpr (MODE == SDNQ_U8) { vmax = fmaxf(vmax, v[i]); vmin = fminf(vmin, v[i]); }
                else amax = fmaxf(amax, fabsf(v[i]));
            }
        }
    }
    float scale, zero = 0.f;
    if constexpr (MODE == SDNQ_U8) {
        vmax = warp_max(vmax);
        vmin = warp_min(vmin);
        if (lane == 0) { s_a[warp] = vmax; s_b[warp] = vmin; }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kWarps; ++i) { vmax = fmaxf(vmax, s_a[i]); vmin = fminf(vmin, s_b[i]); }
        scale = __fdiv_rn(__fsub_rn(vmax, vmin), 255.f);
        zero = __fsub_rn(vmin, __fmul_rn(scale, -128.f));
    } else {
        amax = warp_max(amax);
        if (lane == 0) s_a[warp] = amax;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kWarps; ++i) amax = fmaxf(amax, s_a[i]);
        scale = __fdiv_rn(amax, MODE == SDNQ_F8E4M3 ? 448.f : 127.f);
    }
    const RowDivider divider(scale);
    const bool safe = divider.safe();
    const int d0 = hadamard_dest_dyn(a.hadamard, lane, 0), d1 = hadamard_dest_dyn(a.hadamard, lane, 1);
    int local_sum = 0;
    for (int c = warp; c < chunks; c += kWarps) {
        float v[8];
        fetch(c, v);
        if (c * 256 + lane * 8 >= K) continue;
        const bool want_sum = a.rowsum != nullptr;
        const uint2 r = safe ? quantise8<MODE, true>(v, divider, zero, want_sum, local_sum) : quantise8<MODE, false>(v, divider, zero, want_sum, local_sum);
        const int64_t chunk0 = row * a.K + c * 256;
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
        if (lane == 0) s_sum[warp] = local_sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
#pragma unroll
            for (int i = 0; i < kWarps; ++i) t += s_sum[i];
            a.rowsum[row] = t;
        }
    }
    if (threadIdx.x == 0) {
        a.sx[row] = scale;
        if (a.zx != nullptr) a.zx[row] = zero;
    }
}

template <typename T, bool kConv>
int launch_long(const ActArgs& a, cudaStream_t st) {
    SDNQ_REQUIRE(a.M < (int64_t(1) << 31), SDNQ_EUNSUPPORTED, "act_quant: too many rows");
    const unsigned blocks = static_cast<unsigned>(a.M);
    cudaError_t e;
    if (a.mode == SDNQ_I8) e = launch_pdl(act_quant_long_kernel<T, SDNQ_I8, kConv>, dim3(blocks), dim3(kThreads), 0, st, a);
    else if (a.mode == SDNQ_U8) e = launch_pdl(act_quant_long_kernel<T, SDNQ_U8, kConv>, dim3(blocks), dim3(kThreads), 0, st, a);
    else e = launch_pdl(act_quant_long_kernel<T, SDNQ_F8E4M3, kConv>, dim3(blocks), dim3(kThreads), 0, st, a);
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of act_quant_long_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("act_quant_long_kernel");
}

template <typename T, int WPR, int MAXC, bool kTC, bool kConv>
int launch_mode(const ActArgs& a, cudaStream_t st) {
    constexpr int RPC = kWarps / WPR;
    const int64_t nblk = (a.M + RPC - 1) / RPC;
    // a few waves of resident CTAs walk the row blocks (SDNQ_B200_ACTQ_GRID = CTAs per SM of grid; 0 = one CTA per row block)
    static const int env_per_sm = [] { const char* e = getenv("SDNQ_B200_ACTQ_GRID"); return e != nullptr ? atoi(e) : -1; }();
    const int per_sm = env_per_sm >= 0 ? env_per_sm : (!kConv && sizeof(T) == 2 && MAXC > 4 ? 8 : 32);      // long loops where the next row is prefetched
    const int64_t cap = per_sm > 0 ? int64_t(num_sms()) * per_sm : nblk;
    const unsigned blocks = static_cast<unsigned>(nblk < cap ? nblk : cap);
    cudaError_t e;
    if (a.mode == SDNQ_I8) e = launch_pdl(act_quant_kernel<T, WPR, MAXC, SDNQ_I8, kTC, kConv>, dim3(blocks), dim3(kThreads), 0, st, a);
    else if (a.mode == SDNQ_U8) e = launch_pdl(act_quant_kernel<T, WPR, MAXC, SDNQ_U8, kTC, kConv>, dim3(blocks), dim3(kThreads), 0, st, a);
    else e = launch_pdl(act_quant_kernel<T, WPR, MAXC, SDNQ_F8E4M3, kTC, kConv>, dim3(blocks), dim3(kThreads), 0, st, a);
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of act_quant_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("act_quant_kernel");
}

// SDNQ_B200_HADAMARD_BUTTERFLY=1 keeps the shuffle-butterfly rotation (A/B measurements; f32 activations always use it)
inline bool butterfly_forced() {
    static const bool v = [] {
        const char* e = getenv("SDNQ_B200_HADAMARD_BUTTERFLY");
        return e != nullptr && e[0] != '\0' && e[0] != '0';
    }();
    return v;
}

template <typename T, int WPR, int MAXC, bool kConv>
int launch(const ActArgs& a, cudaStream_t st) {
    if constexpr (sizeof(T) == 2) {
        if (a.hadamard && !butterfly_forced()) return launch_mode<T, WPR, MAXC, true, kConv>(a, st);
    }
    return launch_mode<T, WPR, MAXC, false, kConv>(a, st);
}

template <typename T, bool kConv>
int dispatch(const ActArgs& a, cudaStream_t st) {
    const int64_t chunks = (a.K + 255) / 256;
    // (warps per row, chunks per warp): the unrolled chunk loops of the kernel carry the predicates of every slot, so rows that
    // need 3 (6) chunks per warp -- K = 640 / 1280 / 2560 / 3072 / 5120 (12288): the SD-XL and FLUX widths -- get their own
    // instantiations instead of idling through the 4th (7th, 8th) slot
    if (chunks <= 1) return launch<T, 1, 1, kConv>(a, st);
    if (chunks <= 2) return launch<T, 1, 2, kConv>(a, st);
    if (chunks <= 3) return launch<T, 1, 3, kConv>(a, st);
    if (chunks <= 4) return launch<T, 1, 4, kConv>(a, st);
    if (chunks <= 6) return launch<T, 2, 3, kConv>(a, st);
    if (chunks <= 8) return launch<T, 2, 4, kConv>(a, st);
    if (chunks <= 12) return launch<T, 4, 3, kConv>(a, st);
    if (chunks <= 16) return launch<T, 4, 4, kConv>(a, st);
    if (chunks <= 24) return launch<T, 8, 3, kConv>(a, st);
    if (chunks <= 32) return launch<T, 8, 4, kConv>(a, st);
    if (chunks <= 48) return launch<T, 8, 6, kConv>(a, st);
    if (chunks <= 64) return launch<T, 8, 8, kConv>(a, st);
    return launch_long<T, kConv>(a, st);          // K > 16384: two-pass kernel
}

template <bool kConv>
int act_quant_run(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx, int hadamard_group, int mm_dtype, void* xq,
                         float* sx, float* zx, int32_t* rowsum, void* x_rot, const ConvView& conv, cudaStream_t st) {
    SDNQ_REQUIRE(x && xq && sx, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(M >= 0 && K > 0 && (kConv || ldx >= K), SDNQ_EINVAL, "bad shape M=%lld K=%lld ldx=%lld", (long long)M, (long long)K, (long long)ldx);
    SDNQ_REQUIRE(K % 8 == 0 && (kConv || ldx % 8 == 0), SDNQ_EUNSUPPORTED, "K and ldx must be multiples of 8 (K=%lld ldx=%lld)", (long long)K, (long long)ldx);
    SDNQ_REQUIRE(mm_dtype == SDNQ_I8 || mm_dtype == SDNQ_U8 || mm_dtype == SDNQ_F8E4M3, SDNQ_EINVAL, "bad matmul dtype %d", mm_dtype);
    SDNQ_REQUIRE(mm_dtype != SDNQ_U8 || zx != nullptr, SDNQ_EINVAL, "uint8 activations need a zx output");
    SDNQ_REQUIRE(hadamard_group == 0 || (hadamard_group >= 4 && hadamard_group <= 256 && (hadamard_group & (hadamard_group - 1)) == 0),
                 SDNQ_EUNSUPPORTED, "hadamard group %d: only powers of two in [4,256] are implemented", hadamard_group);
    if (hadamard_group) SDNQ_REQUIRE(K % hadamard_group == 0, SDNQ_EINVAL, "hadamard group %d does not divide K=%lld", hadamard_group, (long long)K);
    SDNQ_REQUIRE(((kConv ? 0 : reinterpret_cast<uintptr_t>(x)) & 15) == 0 && (reinterpret_cast<uintptr_t>(xq) & 7) == 0, SDNQ_EINVAL,
                 "x must be 16-byte and xq 8-byte aligned");
    if (M == 0) return SDNQ_OK;
    float hfac = 1.f;
    if (hadamard_group) {
        hfac = 1.0f / sqrtf(static_cast<float>(hadamard_group));                     // both IEEE-rounded, as on the device
        if (x_dtype == SDNQ_BF16) hfac = __bfloat162float(__float2bfloat16_rn(hfac));
        else if (x_dtype == SDNQ_F16) hfac = __half2float(__float2half_rn(hfac));
    }
    ActArgs a{x, M, K, ldx, hadamard_group, hfac, mm_dtype, reinterpret_cast<uint8_t*>(xq), sx, mm_dtype == SDNQ_U8 ? zx : nullptr, rowsum, x_rot, conv};
    switch (x_dtype) {
        case SDNQ_BF16: return dispatch<__nv_bfloat16, kConv>(a, st);
        case SDNQ_F16: return dispatch<__half, kConv>(a, st);
        case SDNQ_F32: return dispatch<float, kConv>(a, st);
        default: return set_error(SDNQ_EINVAL, "bad activation dtype %d", x_dtype);
    }
}

}  // namespace
}  // namespace sdnq
