// K2 over the im2col view of a convolution input: the same register-resident row quantiser as act_quant.cu, with the row
// gathered straight from the strided NCHW input (see ConvView / conv_gather in act_quant_kernel.cuh).
//
// Reference behaviour restated here:
//   process_conv_input (F.unfold(...).transpose(1, 2))   layers/conv/forward.py:30-76
//   conv_{int8,uint8,fp8}_matmul prologue                layers/conv/conv_int8.py:36-69, conv_uint8.py, conv_fp8.py
#include "act_quant_kernel.cuh"

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
