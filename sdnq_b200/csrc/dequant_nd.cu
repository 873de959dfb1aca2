// Dequant of weights whose scale does not run along contiguous K-groups: convolution and transposed-convolution layers.
//
// Reference behaviour restated here (dequantizer.py:15-84, quantizer.py:95-110, 185-199): a Conv weight [N, C, kh, kw] is
// quantised with the reduction over the input-channel axis only, so its scale / zero-point is [N, 1, kh, kw] (or
// [N, C/g, 1, kh, kw] on the [N, C/g, g, kh, kw] view when grouped), a ConvTranspose weight [C, N, kh, kw] reduces over axis 0
// (scale [1, N, kh, kw], grouped [C, 1, G, kh, kw] on the [C, g, G, kh, kw] view).  dequantize_* is a plain broadcast:
//     W = q * scale            (symmetric)      W = fma(q, scale, zero_point)      (asymmetric, addcmul)
// then `.view(result_shape)`; for SVD layers mm(svd_up, svd_down) (computed in the SVD dtype) is added in f32 before the cast.
//
// One thread owns one octet (8 consecutive values of the flattened quantised view = `bits` storage bytes, exactly as in the
// Linear kernels); its multi-index is found once by mixed-radix division and advanced by carry; the scale index is the dot
// product with the broadcast strides (0 on broadcast axes).  HBM-bound like K3 (bits/8 read + 2 written per element; the
// scale tensor is tiny and lives in L1/L2); these weights are small (a 3x3 conv of SD-XL is <= 14.7 M elements).
#include "unpack.cuh"

namespace sdnq {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxDims = 6;

struct NdArgs {
    const uint8_t* weight;
    const float* scale;
    const float* zp;
    const void* addend;      // optional [numel] tensor added in f32 before the cast (the SVD term), dtype addend_dtype
    int addend_dtype;
    int codebook;
    WFormat f;
    int ndim;
    int64_t octets;
    int32_t dims[kMaxDims];
    int64_t sstride[kMaxDims];
};

__device__ __forceinline__ float addend_at(const void* p, int64_t i, int dtype) {
    if (dtype == SDNQ_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    if (dtype == SDNQ_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
    return reinterpret_cast<const float*>(p)[i];
}

template <int BITS, typename OutT>
__global__ void __launch_bounds__(kThreads) dequant_nd_kernel(const NdArgs a, OutT* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t oct = int64_t(blockIdx.x) * kThreads + threadIdx.x;
    if (oct >= a.octets) return;
    uint32_t codes[8];
    float q[8], w[8];
    octet_values<BITS>(a.weight, oct, a.f, q, codes);
    // multi-index of the first element of the octet
    int idx[kMaxDims];
    int64_t rem = oct * 8, si = 0;
#pragma unroll
    for (int d = kMaxDims - 1; d >= 0; --d) {
        if (d < a.ndim) {
            const int64_t qd = rem / a.dims[d];
            idx[d] = static_cast<int>(rem - qd * a.dims[d]);
            rem = qd;
            si += idx[d] * a.sstride[d];
        } else {
            idx[d] = 0;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (a.codebook) w[i] = a.scale[si * (int64_t(1) << a.f.bits) + codes[i]];
        else if (a.zp != nullptr) w[i] = fmaf(q[i], a.scale[si], a.zp[si]);
        else w[i] = __fmul_rn(q[i], a.scale[si]);
        if (a.addend != nullptr) w[i] = __fadd_rn(w[i], addend_at(a.addend, oct * 8 + i, a.addend_dtype));
        // advance the multi-index by one element (carry from the last axis)
#pragma unroll
        for (int d = kMaxDims - 1; d >= 0; --d) {
            if (d >= a.ndim) continue;
            si += a.sstride[d];
            if (++idx[d] < a.dims[d]) break;
            si -= a.sstride[d] * a.dims[d];
            idx[d] = 0;
        }
    }
    store8<OutT>(out + oct * 8, w);
}

template <int BITS>
int launch_bits(const NdArgs& a, void* out, int out_dtype, cudaStream_t st) {
    const unsigned blocks = static_cast<unsigned>((a.octets + kThreads - 1) / kThreads);
    cudaError_t e;
    if (out_dtype == SDNQ_BF16) e = launch_pdl(dequant_nd_kernel<BITS, __nv_bfloat16>, dim3(blocks), dim3(kThreads), 0, st, a, reinterpret_cast<__nv_bfloat16*>(out));
    else if (out_dtype == SDNQ_F16) e = launch_pdl(dequant_nd_kernel<BITS, __half>, dim3(blocks), dim3(kThreads), 0, st, a, reinterpret_cast<__half*>(out));
    else e = launch_pdl(dequant_nd_kernel<BITS, float>, dim3(blocks), dim3(kThreads), 0, st, a, reinterpret_cast<float*>(out));
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of dequant_nd_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("dequant_nd_kernel");
}

}  // namespace
}  // namespace sdnq

using namespace sdnq;

extern "C" int sdnq_b200_dequant_nd(const void* weight, const sdnq_weight_format* fmt, const float* scale, const float* zero_point, int codebook,
                                    int ndim, const int64_t* dims, const int64_t* scale_strides, const void* addend, int addend_dtype,
                                    void* out, int out_dtype, void* stream) {
    SDNQ_REQUIRE(weight && fmt && scale && out && dims && scale_strides, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(ndim >= 1 && ndim <= kMaxDims, SDNQ_EUNSUPPORTED, "dequant_nd: 1..%d dimensions (got %d)", kMaxDims, ndim);
    SDNQ_REQUIRE(out_dtype == SDNQ_BF16 || out_dtype == SDNQ_F16 || out_dtype == SDNQ_F32, SDNQ_EINVAL, "bad output dtype %d", out_dtype);
    SDNQ_REQUIRE(addend == nullptr || addend_dtype == SDNQ_BF16 || addend_dtype == SDNQ_F16 || addend_dtype == SDNQ_F32, SDNQ_EINVAL, "bad addend dtype %d", addend_dtype);
    WFormat f;
    int rc = make_wformat(fmt, &f);
    if (rc != SDNQ_OK) return rc;
    NdArgs a{};
    int64_t numel = 1;
    for (int d = 0; d < ndim; ++d) {
        SDNQ_REQUIRE(dims[d] > 0 && dims[d] < (int64_t(1) << 31) && scale_strides[d] >= 0, SDNQ_EINVAL, "bad dimension / stride %d", d);
        a.dims[d] = static_cast<int32_t>(dims[d]);
        a.sstride[d] = scale_strides[d];
        numel *= dims[d];
    }
    SDNQ_REQUIRE(numel % 8 == 0, SDNQ_EUNSUPPORTED, "dequant_nd: the number of weights (%lld) must be a multiple of 8", (long long)numel);
    a.weight = reinterpret_cast<const uint8_t*>(weight);
    a.scale = scale;
    a.zp = zero_point;
    a.addend = addend;
    a.addend_dtype = addend_dtype;
    a.codebook = codebook;
    a.f = f;
    a.ndim = ndim;
    a.octets = numel / 8;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (f.bits) {
        case 1: return launch_bits<1>(a, out, out_dtype, st);
        case 2: return launch_bits<2>(a, out, out_dtype, st);
        case 3: return launch_bits<3>(a, out, out_dtype, st);
        case 4: return launch_bits<4>(a, out, out_dtype, st);
        case 5: return launch_bits<5>(a, out, out_dtype, st);
        case 6: return launch_bits<6>(a, out, out_dtype, st);
        case 7: return launch_bits<7>(a, out, out_dtype, st);
        case 8: return launch_bits<8>(a, out, out_dtype, st);
        default: return set_error(SDNQ_EUNSUPPORTED, "dequant_nd: %d-bit weights", f.bits);
    }
}
