// K5p launcher + C ABI: small-M Linear from the stored (packed / grouped) weight.  Kernel body: gemv_packed_kernel.cuh.
#include "gemv_packed_kernel.cuh"

namespace sdnq {
namespace {

template <typename T, int BITS, int MB>
__global__ void __launch_bounds__(gemvp::kThreads) gemv_packed_kernel(const gemvp::Args a) {
    __shared__ float s_red[(gemvp::kWarps - 1) * MB * 4 * 32];
    pdl_launch_dependents();
    pdl_wait();
    gemvp::body<T, BITS, MB>(a, s_red);
}

template <typename T, int BITS>
int launch_mb(const gemvp::Args& a, cudaStream_t st) {
    const int tiles = (a.N + 15) / 16;
    const int cap = num_sms() * 8;
    const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
    const int mb = (a.M + 7) / 8;
    cudaError_t e;
    if (mb <= 1) e = launch_pdl(gemv_packed_kernel<T, BITS, 1>, dim3(grid), dim3(gemvp::kThreads), 0, st, a);
    else if (mb == 2) e = launch_pdl(gemv_packed_kernel<T, BITS, 2>, dim3(grid), dim3(gemvp::kThreads), 0, st, a);
    else e = launch_pdl(gemv_packed_kernel<T, BITS, 4>, dim3(grid), dim3(gemvp::kThreads), 0, st, a);
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of gemv_packed_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("gemv_packed_kernel");
}

}  // namespace
}  // namespace sdnq

using namespace sdnq;

extern "C" int sdnq_b200_linear_small_m_packed(const void* x, int x_dtype, int64_t ldx, const void* weight, const sdnq_weight_format* fmt,
                                               const float* scale, const float* zero_point, int64_t group_size, const void* bias,
                                               int bias_dtype, int64_t bias_ld, void* out, int64_t M, int64_t N, int64_t K, void* stream) {
    SDNQ_REQUIRE(x && weight && scale && out, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(M >= 0 && M <= 32 && N > 0 && K > 0, SDNQ_EINVAL, "small-M Linear: 0 <= M <= 32 (got M=%lld N=%lld K=%lld)", (long long)M, (long long)N, (long long)K);
    SDNQ_REQUIRE(K % 16 == 0 && ldx % 8 == 0 && ldx >= K, SDNQ_EUNSUPPORTED, "small-M Linear: K %% 16 == 0 and ldx %% 8 == 0 (K=%lld ldx=%lld)", (long long)K, (long long)ldx);
    SDNQ_REQUIRE(x_dtype == SDNQ_BF16 || x_dtype == SDNQ_F16, SDNQ_EUNSUPPORTED, "small-M Linear: bf16 / f16 activations (got %d)", x_dtype);
    SDNQ_REQUIRE(bias == nullptr || bias_dtype == SDNQ_BF16 || bias_dtype == SDNQ_F16 || bias_dtype == SDNQ_F32, SDNQ_EINVAL, "bad bias dtype %d", bias_dtype);
    SDNQ_REQUIRE(bias_ld == 0 || bias_ld >= N, SDNQ_EINVAL, "bias_ld must be 0 (vector bias) or >= N");
    WFormat f;
    int rc = make_wformat(fmt, &f);
    if (rc != SDNQ_OK) return rc;
    SDNQ_REQUIRE(f.word_bytes == 1, SDNQ_EUNSUPPORTED, "small-M Linear: 1-bit weights stored as int64 words are not covered");
    const int64_t group = (group_size <= 0 || group_size >= K) ? K : group_size;
    SDNQ_REQUIRE(group % 8 == 0 && K % group == 0, SDNQ_EUNSUPPORTED, "small-M Linear: group size must be a multiple of 8 dividing K (group=%lld K=%lld)",
                 (long long)group, (long long)K);
    SDNQ_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0, SDNQ_EINVAL, "x and weight must be 16-byte aligned");
    SDNQ_REQUIRE(N * K < (int64_t(1) << 40) && N < (int64_t(1) << 31) && K < (int64_t(1) << 31), SDNQ_EUNSUPPORTED, "weight too large");
    if (M == 0) return SDNQ_OK;
    gemvp::Args a{x, ldx, reinterpret_cast<const uint8_t*>(weight), scale, zero_point, bias, bias_dtype, bias_ld, out,
                  int(M), int(N), int(K), int(group), int(K / group), f};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (x_dtype == SDNQ_BF16) { SDNQ_DISPATCH_BITS(f.bits, return (launch_mb<__nv_bfloat16, BITS>(a, st))); }
    else { SDNQ_DISPATCH_BITS(f.bits, return (launch_mb<__half, BITS>(a, st))); }
    return SDNQ_OK;
}
