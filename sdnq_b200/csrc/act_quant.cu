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
#include "act_quant_kernel.cuh"

namespace sdnq {

int act_quant_impl(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx, int hadamard_group, int mm_dtype, void* xq,
                   float* sx, float* zx, int32_t* rowsum, void* x_rot, cudaStream_t st) {
    ConvView none{};
    return act_quant_run<false>(x, x_dtype, M, K, ldx, hadamard_group, mm_dtype, xq, sx, zx, rowsum, x_rot, none, st);
}

}  // namespace sdnq


extern "C" int sdnq_b200_act_quant(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx, int hadamard_group, int mm_dtype,
                                   void* xq, float* sx, float* zx, int32_t* rowsum, void* x_rot, void* stream) {
    return sdnq::act_quant_impl(x, x_dtype, M, K, ldx, hadamard_group, mm_dtype, xq, sx, zx, rowsum, x_rot,
                                reinterpret_cast<cudaStream_t>(stream));
}
