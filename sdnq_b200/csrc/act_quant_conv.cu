// K2 over the im2col view of a convolution input: the same register-resident row quantiser as act_quant.cu, with the row
// gathered straight from the strided NCHW input (see ConvView / conv_gather in act_quant_kernel.cuh).
//
// Reference behaviour restated here:
//   process_conv_input (F.unfold(...).transpose(1, 2))   layers/conv/forward.py:30-76
//   conv_{int8,uint8,fp8}_matmul prologue                layers/conv/conv_int8.py:36-69, conv_uint8.py, conv_fp8.py
#include "act_quant_kernel.cuh"

#include <algorithm>

namespace sdnq {

int conv_act_quant_impl(const void* x, int x_dtype, const sdnq_conv2d_geometry* g, int hadamard_group, int mm_dtype, void* xq,
                        float* sx, float* zx, int32_t* rowsum, void* x_rot, cudaStream_t st) {
    SDNQ_REQUIRE(g != nullptr, SDNQ_EINVAL, "NULL geometry");
    SDNQ_REQUIRE(g->batch >= 0 && g->channels > 0 && g->height > 0 && g->width > 0 && g->kernel_h > 0 && g->kernel_w > 0 && g->stride_h > 0 &&
                 g->stride_w > 0 && g->dilation_h > 0 && g->dilation_w > 0 && g->pad_h >= 0 && g->pad_w >= 0, SDNQ_EINVAL, "bad convolution geometry");
    const int64_t Hout = (g->height + 2 * g->pad_h - g->dilation_h * (g->kernel_h - 1) - 1) / g->stride_h + 1;
    const int64_t Wout = (g->width + 2 * g->pad_w - g->dilation_w * (g->kernel_w - 1) - 1) / g->stride_w + 1;
    SDNQ_REQUIRE(Hout > 0 && Wout > 0, SDNQ_EINVAL, "empty convolution output (%lld x %lld)", (long long)Hout, (long long)Wout);
    const int64_t M = g->batch * Hout * Wout, K = g->channels * g->kernel_h * g->kernel_w;
    SDNQ_REQUIRE(M < (int64_t(1) << 31) && Hout * Wout < (int64_t(1) << 31) && K < (int64_t(1) << 31), SDNQ_EUNSUPPORTED, "convolution too large");
    // per-image offsets are computed in 32 bits on the device
    const int64_t span = (g->channels - 1) * (g->x_stride_c < 0 ? -g->x_stride_c : g->x_stride_c) + (g->height - 1) * (g->x_stride_h < 0 ? -g->x_stride_h : g->x_stride_h) +
                         (g->width - 1) * (g->x_stride_w < 0 ? -g->x_stride_w : g->x_stride_w);
    SDNQ_REQUIRE(span < (int64_t(1) << 31) && g->x_stride_c >= 0 && g->x_stride_h >= 0 && g->x_stride_w >= 0, SDNQ_EUNSUPPORTED, "input image too large or negatively strided");
    ConvView cv{1, int(g->channels), int(g->height), int(g->width), int(g->kernel_h), int(g->kernel_w), int(g->stride_h), int(g->stride_w),
                int(g->pad_h), int(g->pad_w), int(g->dilation_h), int(g->dilation_w), int(Wout), int(Hout * Wout),
                g->x_stride_b, g->x_stride_c, g->x_stride_h, g->x_stride_w};
    return act_quant_run<true>(x, x_dtype, M, K, K, hadamard_group, mm_dtype, xq, sx, zx, rowsum, x_rot, cv, st);
}

}  // namespace sdnq

extern "C" int sdnq_b200_conv_act_quant(const void* x, int x_dtype, const sdnq_conv2d_geometry* geometry, int hadamard_group, int mm_dtype,
                                        void* xq, float* sx, float* zx, int32_t* rowsum, void* x_rot, void* stream) {
    return sdnq::conv_act_quant_impl(x, x_dtype, geometry, hadamard_group, mm_dtype, xq, sx, zx, rowsum, x_rot,
                                     reinterpret_cast<cudaStream_t>(stream));
}

// =====================================================================================================================
// Tiled im2col quantiser (the fast path for 1x1 and 3x3 kernels without rotation): coalesced in both directions.
//
// The gather kernel above reads each im2col row with scalar 2-byte loads that do not coalesce (a row walks (c, i, j) with the
// image row / plane strides between consecutive elements): 0.16-0.22 TB/s.  Here
//   pass A  conv_pixel_stats_kernel : per *input* pixel, the maximum |x| (or min / max) over the channels -- one coalesced
//           read of x; the statistic of an im2col row is then the maximum over the taps of its window (exact: max is exact),
//   pass B  conv_quant_tile_kernel  : a CTA owns 32 consecutive output pixels (lane = pixel, so for a fixed (c, i, j) a warp
//           reads 32 neighbouring input pixels: coalesced, and the 9 taps hit the same cache lines) and walks the channels in
//           chunks of 32 (warp w: channels 4w .. 4w+3 of the chunk); codes are packed four channels x taps at a time into a
//           shared-memory tile [pixel][chunk columns] (odd word pitch: conflict-free both ways) and written out as 4-byte
//           stores that cover each row segment contiguously.
// Same arithmetic as everywhere else (scale = amax / 127 | 448, correctly rounded division, round-half-even / e4m3 RNE), so the
// codes, scales, zero points and row sums are bit-identical to the gather kernel and to the reference's unfold + quantise.
namespace sdnq {
namespace {

constexpr int kTilePixels = 32;
constexpr int kChunkCh = 32;          // channels per chunk = 8 warps x 4

struct ConvTileArgs {
    const void* x;
    ConvView cv;
    int64_t M;                        // output pixels
    int K;                            // C * taps
    float* stats;                     // [B*H*W] amax, or [B*H*W][2] min / max (uint8 mode)
    uint8_t* xq;
    float* sx;
    float* zx;
    int32_t* rowsum;
};

template <typename T, bool kMinMax>
__global__ void __launch_bounds__(256) conv_pixel_stats_kernel(const ConvTileArgs a, int64_t pixels) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t p = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (p >= pixels) return;
    const int hw = a.cv.H * a.cv.W;
    const int b = static_cast<int>(p / hw);
    const int r = static_cast<int>(p - int64_t(b) * hw);
    const int h = r / a.cv.W, w = r - h * a.cv.W;
    const T* px = reinterpret_cast<const T*>(a.x) + int64_t(b) * a.cv.sB + int64_t(h) * a.cv.sH + int64_t(w) * a.cv.sW;
    float amax = 0.f, vmin = INFINITY, vmax = -INFINITY;
    // blockIdx.y splits the channels (small images have too few pixels to fill the machine): partial maxima are combined with
    // atomicMax on the bit pattern (non-negative floats order like unsigned integers; the buffer is zeroed by the launcher)
    const int per = (a.cv.C + int(gridDim.y) - 1) / int(gridDim.y);
    const int c_begin = int(blockIdx.y) * per, c_end = min(a.cv.C, c_begin + per);
#pragma unroll 8
    for (int c = c_begin; c < c_end; ++c) {
        const float v = ElemTraits<T>::load(px[int64_t(c) * a.cv.sC]);
        if constexpr (kMinMax) { vmin = fminf(vmin, v); vmax = fmaxf(vmax, v); }
        else amax = fmaxf(amax, fabsf(v));
    }
    if constexpr (kMinMax) { a.stats[2 * p] = vmin; a.stats[2 * p + 1] = vmax; }      // launched with gridDim.y == 1
    else if (gridDim.y == 1) a.stats[p] = amax;
    else atomicMax(reinterpret_cast<unsigned int*>(a.stats) + p, __float_as_uint(amax));
}

// Four consecutive channels x TAPS window elements of one output pixel -> TAPS packed words (column order (c, i, j)) in the
// shared-memory tile; returns the sum of the integer codes.
template <typename T, int MODE, int TAPS, bool kSafe>
__device__ __forceinline__ int quantise_channels(const T* __restrict__ xc, int64_t sC, const int (&toff)[TAPS], uint32_t tmask,
                                                 const actq::RowDivider& divider, float zero, bool want_sum, uint32_t* __restrict__ dst) {
    float v[4 * TAPS];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int tp = 0; tp < TAPS; ++tp) v[q * TAPS + tp] = (tmask >> tp) & 1u ? ElemTraits<T>::load(xc[q * sC + toff[tp]]) : 0.f;
    }
    int sum = 0;
#pragma unroll
    for (int wd = 0; wd < TAPS; ++wd) {                       // word wd = columns 4wd .. 4wd+3 of the four-channel run
        float qv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float x0 = v[4 * wd + i];
            if constexpr (MODE == SDNQ_U8) x0 = __fsub_rn(x0, zero);
            qv[i] = divider.template div<kSafe>(x0);
        }
        if constexpr (MODE == SDNQ_F8E4M3) {
            if constexpr (!kSafe) {                           // nan_to_num (0/0 on an all-zero row)
#pragma unroll
                for (int i = 0; i < 4; ++i) if (qv[i] != qv[i]) qv[i] = 0.f;
            }
            // cvt.rn.satfinite.e4m3x2 saturates to +-448 = the reference's clamp before the cast
            const uint32_t lo = static_cast<uint16_t>(__nv_cvt_float2_to_fp8x2(make_float2(qv[0], qv[1]), __NV_SATFINITE, __NV_E4M3));
            const uint32_t hi = static_cast<uint16_t>(__nv_cvt_float2_to_fp8x2(make_float2(qv[2], qv[3]), __NV_SATFINITE, __NV_E4M3));
            dst[wd] = lo | (hi << 16);
        } else {
            int c[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) c[i] = __float2int_rn(qv[i]);      // round-half-even; NaN -> 0
            uint32_t w;
#ifdef SDNQ_HOST_EMU
            w = ::sdnq_emu::pack_sat_s8x4(c[0], c[1], c[2], c[3]);
#else
            // cvt.pack.sat.s8.s32.b32 d, a, b, c :  d = (c << 16) | (sat8(a) << 8) | sat8(b)
            asm("{\n\t.reg .b32 t;\n\tcvt.pack.sat.s8.s32.b32 t, %4, %3, 0;\n\tcvt.pack.sat.s8.s32.b32 %0, %2, %1, t;\n\t}"
                : "=r"(w) : "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]));
#endif
            dst[wd] = w;
            if (want_sum) {
#pragma unroll
                for (int i = 0; i < 4; ++i) sum += max(-128, min(127, c[i]));
            }
        }
    }
    return sum;
}

template <typename T, int MODE, int KH, int KW>
__global__ void __launch_bounds__(256, 3) conv_quant_tile_kernel(const ConvTileArgs a) {
    constexpr int TAPS = KH * KW;
    constexpr int kPitch = kChunkCh * TAPS / 4 + 1;                  // words per tile row (odd)
    __shared__ uint32_t s_tiles[2][kTilePixels * kPitch];            // double-buffered: one barrier per chunk
    __shared__ int s_sum[8][kTilePixels];
    pdl_launch_dependents();
    pdl_wait();
    using namespace actq;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const ConvView& cv = a.cv;
    const int64_t m = int64_t(blockIdx.x) * kTilePixels + lane;
    const bool m_ok = m < a.M;
    int ih0 = 0, iw0 = 0;
    int64_t img = 0, stat0 = 0;
    {
        const int64_t mm = m_ok ? m : 0;
        const int b = static_cast<int>(mm / cv.HWout);
        const int rem = static_cast<int>(mm - int64_t(b) * cv.HWout);
        const int oh = rem / cv.Wout, ow = rem - oh * cv.Wout;
        ih0 = oh * cv.sh - cv.ph;
        iw0 = ow * cv.sw - cv.pw;
        img = int64_t(b) * cv.sB;
        stat0 = int64_t(b) * cv.H * cv.W;
    }
    // ---- the row statistic from the per-pixel statistics of the window (every warp computes its own copy: 9 cached loads)
    float amax = 0.f, vmin = INFINITY, vmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < KH; ++i) {
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            const int ih = ih0 + i * cv.dh, iw = iw0 + j * cv.dw;
            const bool in = m_ok && ih >= 0 && ih < cv.H && iw >= 0 && iw < cv.W;
            const int64_t sp = stat0 + int64_t(ih) * cv.W + iw;
            if constexpr (MODE == SDNQ_U8) {
                const float lo = in ? a.stats[2 * sp] : 0.f, hi = in ? a.stats[2 * sp + 1] : 0.f;      // padding reads as 0
                vmin = fminf(vmin, lo);
                vmax = fmaxf(vmax, hi);
            } else {
                if (in) amax = fmaxf(amax, a.stats[sp]);
            }
        }
    }
    float scale, zero = 0.f;
    if constexpr (MODE == SDNQ_U8) {
        scale = __fdiv_rn(__fsub_rn(vmax, vmin), 255.f);
        zero = __fsub_rn(vmin, __fmul_rn(scale, -128.f));
    } else {
        scale = __fdiv_rn(amax, MODE == SDNQ_F8E4M3 ? 448.f : 127.f);
    }
    const RowDivider divider(scale);
    const bool safe = divider.safe();
    if (warp == 0 && m_ok && blockIdx.y == 0) {
        a.sx[m] = scale;
        if (a.zx != nullptr) a.zx[m] = zero;
    }
    const T* xim = reinterpret_cast<const T*>(a.x) + img;
    // tap offsets / validity of this pixel's window (registers: TAPS ints + a bit mask)
    int toff[TAPS];
    uint32_t tmask = 0;
#pragma unroll
    for (int i = 0; i < KH; ++i) {
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            const int ih = ih0 + i * cv.dh, iw = iw0 + j * cv.dw;
            const bool in = m_ok && ih >= 0 && ih < cv.H && iw >= 0 && iw < cv.W;
            toff[i * KW + j] = in ? static_cast<int>(int64_t(ih) * cv.sH + int64_t(iw) * cv.sW) : 0;
            tmask |= in ? (1u << (i * KW + j)) : 0u;
        }
    }
    int local_sum = 0;
    const bool want_sum = a.rowsum != nullptr;
    uint8_t* const out_row0 = a.xq + int64_t(blockIdx.x) * kTilePixels * a.K;
    // blockIdx.y splits the channel chunks of a tile over several CTAs when there are too few tiles to fill the machine
    int buf = 0;
    for (int c0 = int(blockIdx.y) * kChunkCh; c0 < cv.C; c0 += int(gridDim.y) * kChunkCh, buf ^= 1) {
        uint32_t* const s_tile = s_tiles[buf];
        const int cw = c0 + 4 * warp;                                  // this warp's four channels
        if (cw < cv.C) {                                               // C % 4 == 0 (host-checked)
            uint32_t* dst = s_tile + lane * kPitch + warp * TAPS;
            const T* xc = xim + int64_t(cw) * cv.sC;
            // one uniform branch per chunk (not per element) on the division path, as in quantise8
            if (safe) local_sum += quantise_channels<T, MODE, TAPS, true>(xc, cv.sC, toff, tmask, divider, zero, want_sum, dst);
            else local_sum += quantise_channels<T, MODE, TAPS, false>(xc, cv.sC, toff, tmask, divider, zero, want_sum, dst);
        }
        __syncthreads();
        // ---- write the tile out: row p holds (chunk channels) * TAPS bytes, contiguous in xq at column c0 * TAPS
        const int chunk_ch = min(kChunkCh, cv.C - c0);
        const int row_words = chunk_ch * TAPS / 4;
        const int64_t rows_left = a.M - int64_t(blockIdx.x) * kTilePixels;
        const int rows = rows_left < kTilePixels ? static_cast<int>(rows_left) : kTilePixels;
        for (int p = warp; p < rows; p += 8) {                         // a warp writes whole row segments: no index division
            uint8_t* orow = out_row0 + int64_t(p) * a.K + c0 * TAPS;
            for (int col = lane; col < row_words; col += 32) *reinterpret_cast<uint32_t*>(orow + 4 * col) = s_tile[p * kPitch + col];
        }
        // no second barrier: the next chunk fills the other buffer, and this one is refilled only after the next chunk's barrier,
        // which every thread reaches after finishing the stores above
    }
    if (want_sum) {
        s_sum[warp][lane] = local_sum;
        __syncthreads();
        if (warp == 0 && m_ok) {
            int t = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += s_sum[w][lane];
            if (gridDim.y == 1) a.rowsum[m] = t;
            else atomicAdd(a.rowsum + m, t);                            // zeroed by the launcher
        }
    }
}

template <typename T, int KH, int KW>
int launch_tiled(const ConvTileArgs& a, int mode, int64_t in_pixels, cudaStream_t st) {
    const unsigned gs = static_cast<unsigned>((in_pixels + 255) / 256), gt = static_cast<unsigned>((a.M + kTilePixels - 1) / kTilePixels);
    const int target = num_sms() * 3;
    // channel splits: enough CTAs to fill the machine when the image is small (many channels, few pixels)
    unsigned ss = 1, ts = 1;
    if (mode != SDNQ_U8 && int(gs) < target) ss = static_cast<unsigned>(std::min<int>(std::max(1, a.cv.C / 32), (target + int(gs) - 1) / int(gs)));
    const int chunks = (a.cv.C + kChunkCh - 1) / kChunkCh;
    // ... and about four CTAs per resident slot, so that the last wave of tiles does not leave most of the machine idle
    if (int(gt) < 4 * target) ts = static_cast<unsigned>(std::min<int>(chunks, (4 * target + int(gt) - 1) / int(gt)));
    if (ss > 1) SDNQ_CUDA_OK(cudaMemsetAsync(a.stats, 0, sizeof(float) * in_pixels, st));
    if (ts > 1 && a.rowsum != nullptr) SDNQ_CUDA_OK(cudaMemsetAsync(a.rowsum, 0, sizeof(int32_t) * a.M, st));
    cudaError_t e = mode == SDNQ_U8 ? launch_pdl(conv_pixel_stats_kernel<T, true>, dim3(gs, 1), dim3(256), 0, st, a, in_pixels)
                                    : launch_pdl(conv_pixel_stats_kernel<T, false>, dim3(gs, ss), dim3(256), 0, st, a, in_pixels);
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of conv_pixel_stats_kernel failed: %s", cudaGetErrorString(e));
    count_launch();
    if (mode == SDNQ_I8) e = launch_pdl(conv_quant_tile_kernel<T, SDNQ_I8, KH, KW>, dim3(gt, ts), dim3(256), 0, st, a);
    else if (mode == SDNQ_U8) e = launch_pdl(conv_quant_tile_kernel<T, SDNQ_U8, KH, KW>, dim3(gt, ts), dim3(256), 0, st, a);
    else e = launch_pdl(conv_quant_tile_kernel<T, SDNQ_F8E4M3, KH, KW>, dim3(gt, ts), dim3(256), 0, st, a);
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of conv_quant_tile_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("conv_quant_tile_kernel");
}

template <typename T>
int dispatch_tiled(const ConvTileArgs& a, int mode, int64_t in_pixels, cudaStream_t st) {
    if (a.cv.kh == 1 && a.cv.kw == 1) return launch_tiled<T, 1, 1>(a, mode, in_pixels, st);
    return launch_tiled<T, 3, 3>(a, mode, in_pixels, st);
}

}  // namespace
}  // namespace sdnq

extern "C" size_t sdnq_b200_conv_act_quant_workspace_bytes(const sdnq_conv2d_geometry* g, int mm_dtype) {
    if (g == nullptr || g->batch <= 0) return 0;
    return static_cast<size_t>(g->batch) * g->height * g->width * sizeof(float) * (mm_dtype == SDNQ_U8 ? 2 : 1);
}

extern "C" int sdnq_b200_conv_act_quant_ws(const void* x, int x_dtype, const sdnq_conv2d_geometry* g, int hadamard_group, int mm_dtype,
                                           void* xq, float* sx, float* zx, int32_t* rowsum, void* x_rot, void* workspace,
                                           size_t workspace_bytes, void* stream) {
    using namespace sdnq;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool tiled_kernel = g != nullptr && ((g->kernel_h == 1 && g->kernel_w == 1) || (g->kernel_h == 3 && g->kernel_w == 3));
    const char* env = getenv("SDNQ_B200_CONV_GATHER");              // "1": always use the gather kernel (A/B measurements)
    if (!tiled_kernel || hadamard_group != 0 || x_rot != nullptr || workspace == nullptr || (env != nullptr && env[0] == '1') ||
        g->channels % 4 != 0 || (g->channels * g->kernel_h * g->kernel_w) % 4 != 0 || (reinterpret_cast<uintptr_t>(xq) & 3) != 0)
        return conv_act_quant_impl(x, x_dtype, g, hadamard_group, mm_dtype, xq, sx, zx, rowsum, x_rot, st);
    SDNQ_REQUIRE(x && xq && sx, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(g->batch >= 0 && g->channels > 0 && g->height > 0 && g->width > 0 && g->stride_h > 0 && g->stride_w > 0 && g->dilation_h > 0 &&
                 g->dilation_w > 0 && g->pad_h >= 0 && g->pad_w >= 0, SDNQ_EINVAL, "bad convolution geometry");
    SDNQ_REQUIRE(mm_dtype == SDNQ_I8 || mm_dtype == SDNQ_U8 || mm_dtype == SDNQ_F8E4M3, SDNQ_EINVAL, "bad matmul dtype %d", mm_dtype);
    SDNQ_REQUIRE(mm_dtype != SDNQ_U8 || zx != nullptr, SDNQ_EINVAL, "uint8 activations need a zx output");
    SDNQ_REQUIRE(workspace_bytes >= sdnq_b200_conv_act_quant_workspace_bytes(g, mm_dtype) && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0,
                 SDNQ_EINVAL, "conv workspace too small or misaligned");
    const int64_t Hout = (g->height + 2 * g->pad_h - g->dilation_h * (g->kernel_h - 1) - 1) / g->stride_h + 1;
    const int64_t Wout = (g->width + 2 * g->pad_w - g->dilation_w * (g->kernel_w - 1) - 1) / g->stride_w + 1;
    SDNQ_REQUIRE(Hout > 0 && Wout > 0, SDNQ_EINVAL, "empty convolution output");
    const int64_t M = g->batch * Hout * Wout, K = g->channels * g->kernel_h * g->kernel_w;
    SDNQ_REQUIRE(M < (int64_t(1) << 31) && K < (int64_t(1) << 31) && g->batch * g->height * g->width < (int64_t(1) << 31), SDNQ_EUNSUPPORTED, "convolution too large");
    const int64_t span = (g->channels - 1) * g->x_stride_c + (g->height - 1) * g->x_stride_h + (g->width - 1) * g->x_stride_w;
    SDNQ_REQUIRE(g->x_stride_c >= 0 && g->x_stride_h >= 0 && g->x_stride_w >= 0 && span < (int64_t(1) << 31), SDNQ_EUNSUPPORTED, "input image too large or negatively strided");
    if (M == 0) return SDNQ_OK;
    ConvView cv{1, int(g->channels), int(g->height), int(g->width), int(g->kernel_h), int(g->kernel_w), int(g->stride_h), int(g->stride_w),
                int(g->pad_h), int(g->pad_w), int(g->dilation_h), int(g->dilation_w), int(Wout), int(Hout * Wout),
                g->x_stride_b, g->x_stride_c, g->x_stride_h, g->x_stride_w};
    ConvTileArgs a{x, cv, M, int(K), reinterpret_cast<float*>(workspace), reinterpret_cast<uint8_t*>(xq), sx, mm_dtype == SDNQ_U8 ? zx : nullptr, rowsum};
    const int64_t in_pixels = g->batch * g->height * g->width;
    switch (x_dtype) {
        case SDNQ_BF16: return dispatch_tiled<__nv_bfloat16>(a, mm_dtype, in_pixels, st);
        case SDNQ_F16: return dispatch_tiled<__half>(a, mm_dtype, in_pixels, st);
        case SDNQ_F32: return dispatch_tiled<float>(a, mm_dtype, in_pixels, st);
        default: return set_error(SDNQ_EINVAL, "bad activation dtype %d", x_dtype);
    }
}

// =====================================================================================================================
// [B*HW, N] GEMM output -> NCHW: the `.view(B, H_out, W_out, N).permute(0, 3, 1, 2).contiguous()` at the end of
// conv_int8_matmul (layers/conv/conv_int8.py:83-89) as one shared-memory tiled transpose (64 x 64 elements per CTA, both
// sides coalesced) instead of a generic strided copy.
namespace sdnq {
namespace {

template <typename E>
__global__ void __launch_bounds__(256) rows_to_nchw_kernel(const E* __restrict__ in, E* __restrict__ out, int HW, int N) {
    __shared__ E tile[64][64 + 2];
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.z, p0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;            // 64 x 4
    const E* src = in + (int64_t(b) * HW + p0) * N + n0;
#pragma unroll
    for (int r = ty; r < 64; r += 4)
        if (p0 + r < HW && n0 + tx < N) tile[r][tx] = src[int64_t(r) * N + tx];
    __syncthreads();
    E* dst = out + (int64_t(b) * N + n0) * HW + p0;
#pragma unroll
    for (int r = ty; r < 64; r += 4)
        if (n0 + r < N && p0 + tx < HW) dst[int64_t(r) * HW + tx] = tile[tx][r];
}

}  // namespace
}  // namespace sdnq

extern "C" int sdnq_b200_rows_to_nchw(const void* in, void* out, int elem_bytes, int64_t batch, int64_t hw, int64_t channels, void* stream) {
    using namespace sdnq;
    SDNQ_REQUIRE(in && out, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(elem_bytes == 2 || elem_bytes == 4, SDNQ_EINVAL, "element size must be 2 or 4 bytes (got %d)", elem_bytes);
    SDNQ_REQUIRE(batch >= 0 && batch < 65536 && hw > 0 && channels > 0 && hw < (int64_t(1) << 31) && channels < (int64_t(1) << 22), SDNQ_EINVAL,
                 "bad shape batch=%lld hw=%lld channels=%lld", (long long)batch, (long long)hw, (long long)channels);
    if (batch == 0) return SDNQ_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const dim3 grid(static_cast<unsigned>((hw + 63) / 64), static_cast<unsigned>((channels + 63) / 64), static_cast<unsigned>(batch));
    cudaError_t e = elem_bytes == 2 ? launch_pdl(rows_to_nchw_kernel<uint16_t>, grid, dim3(256), 0, st, reinterpret_cast<const uint16_t*>(in),
                                                 reinterpret_cast<uint16_t*>(out), int(hw), int(channels))
                                    : launch_pdl(rows_to_nchw_kernel<uint32_t>, grid, dim3(256), 0, st, reinterpret_cast<const uint32_t*>(in),
                                                 reinterpret_cast<uint32_t*>(out), int(hw), int(channels));
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of rows_to_nchw_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("rows_to_nchw_kernel");
}
