// Library-level pieces of the C ABI: error reporting, device checks, launch accounting and the fused
// W8A8 Linear entry point (K2 + K1 on one stream with a caller-owned workspace).
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace sdnq {

namespace {
thread_local char g_error[512] = "";
thread_local int64_t g_launches = 0;
}  // namespace

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches += n; }

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("SDNQ_B200_PDL");
        return !(e && (e[0] == '0' || e[0] == 'n' || e[0] == 'N' || e[0] == 'f' || e[0] == 'F'));
    }();
    return on;
}

int num_sms() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return cached;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) cached = v;
        cached_dev = dev;
    }
    return cached;
}

// implemented in act_quant.cu / gemm_w8a8.cu
int act_quant_impl(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx, int hadamard_group, int mm_dtype, void* xq,
                   float* sx, float* zx, int32_t* rowsum, void* x_rot, cudaStream_t st);

}  // namespace sdnq

using namespace sdnq;

extern "C" int sdnq_b200_abi_version(void) { return SDNQ_B200_ABI_VERSION; }

extern "C" const char* sdnq_b200_last_error(void) { return g_error; }

extern "C" int sdnq_b200_check_device(int device) {
    int major = 0, minor = 0;
    SDNQ_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    SDNQ_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    SDNQ_REQUIRE(major == 10, SDNQ_EARCH, "device %d is sm_%d%d; these kernels are built for sm_100a only", device, major, minor);
    return SDNQ_OK;
}

extern "C" int64_t sdnq_b200_launch_count(int reset) {
    const int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

// workspace layout: xq [M*K] | sx [M] f32 | zx [M] f32 | rowsum [M] i32   (each section 256 B aligned)
static inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

extern "C" size_t sdnq_b200_linear_w8a8_workspace_bytes(int64_t M, int64_t K) {
    if (M <= 0 || K <= 0) return 0;
    return align256(size_t(M) * size_t(K)) + 3 * align256(size_t(M) * 4);
}

extern "C" int sdnq_b200_linear_w8a8(const void* x, int x_dtype, int64_t ldx, const void* wq, int mm_dtype, const float* sw,
                                     const float* zp, const int32_t* colsum, const void* bias, int bias_dtype, int hadamard_group,
                                     void* out, int out_dtype, int64_t M, int64_t N, int64_t K, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    SDNQ_REQUIRE(mm_dtype == SDNQ_I8 || mm_dtype == SDNQ_U8 || mm_dtype == SDNQ_F8E4M3, SDNQ_EINVAL, "bad matmul dtype %d", mm_dtype);
    SDNQ_REQUIRE(workspace != nullptr && workspace_bytes >= sdnq_b200_linear_w8a8_workspace_bytes(M, K), SDNQ_EINVAL,
                 "workspace too small: %zu < %zu", workspace_bytes, sdnq_b200_linear_w8a8_workspace_bytes(M, K));
    SDNQ_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, SDNQ_EINVAL, "workspace must be 256-byte aligned");
    SDNQ_REQUIRE(mm_dtype != SDNQ_U8 || colsum != nullptr, SDNQ_EINVAL, "uint8 matmul needs the weight column sums");
    if (M == 0) return SDNQ_OK;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    void* xq = ws;
    float* sx = reinterpret_cast<float*>(ws + align256(size_t(M) * size_t(K)));
    float* zx = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sx) + align256(size_t(M) * 4));
    int32_t* rowsum = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(zx) + align256(size_t(M) * 4));
    const bool need_rowsum = zp != nullptr;
    int rc = sdnq_b200_act_quant(x, x_dtype, M, K, ldx, hadamard_group, mm_dtype, xq, sx, mm_dtype == SDNQ_U8 ? zx : nullptr,
                                 need_rowsum ? rowsum : nullptr, nullptr, stream);
    if (rc != SDNQ_OK) return rc;
    return sdnq_b200_scaled_mm(xq, wq, mm_dtype == SDNQ_F8E4M3 ? SDNQ_F8E4M3 : SDNQ_I8, sx, sw, bias, bias_dtype, 0,
                               need_rowsum ? rowsum : nullptr, zp, mm_dtype == SDNQ_U8 ? colsum : nullptr,
                               mm_dtype == SDNQ_U8 ? zx : nullptr, out, out_dtype, M, N, K, stream);
}
