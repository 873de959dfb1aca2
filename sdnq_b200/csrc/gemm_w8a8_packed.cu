// K1 instantiations for packed weights of any width: the stored 2 / 3 / 5 / 6 / 7-bit codes (packed_int/pack.py octets) are staged by
// TMA and expanded tile by tile into the UMMA operand ring by the GEMM's unpack warps -- integer codes to int8 (kind::i8), minifloat
// codes to the e4m3 byte of the same value (kind::f8f6f4) -- so the layer keeps no expanded N x K copy of its weight
// (reference: unpack + transpose on every call, layers/linear/linear_int8.py:38-44, linear_fp8.py:38, packed_int/unpack.py:233-372,
// packed_float.py:85-132).  A separate translation unit so that these instantiations compile next to gemm_w8a8.cu.
#include "gemm_w8a8_kernel.cuh"

namespace sdnq {

int launch_gemm_packed_any(const void* a, const void* b, const GemmParams& p, cudaStream_t st) {
    const bool i8 = p.pk_kind == 0;
    switch (p.out_dtype) {
        case SDNQ_BF16: return i8 ? launch_gemm<128, true, OUT_BF16, false, 7>(a, b, p, st) : launch_gemm<128, false, OUT_BF16, false, 7>(a, b, p, st);
        case SDNQ_F16: return i8 ? launch_gemm<128, true, OUT_F16, false, 7>(a, b, p, st) : launch_gemm<128, false, OUT_F16, false, 7>(a, b, p, st);
        default: return i8 ? launch_gemm<128, true, OUT_F32, false, 7>(a, b, p, st) : launch_gemm<128, false, OUT_F32, false, 7>(a, b, p, st);
    }
}

}  // namespace sdnq
