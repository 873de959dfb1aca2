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
int linear_fused_impl(const void* x, int x_dtype, int64_t ldx, const void* wq, int ab_dtype, const float* sw, const void* bias,
                      int bias_dtype, void* out, int out_dtype, int64_t M, int64_t N, int64_t K, uint8_t* xq, float* sx, int* sync,
                      cudaStream_t st);

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

extern "C" int64_t sdnq_b200_stream_capture_id(void* stream) {
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    unsigned long long id = 0;
    if (cudaStreamGetCaptureInfo(reinterpret_cast<cudaStream_t>(stream), &status, &id) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return status == cudaStreamCaptureStatusActive ? static_cast<int64_t>(id) : 0;
}

extern "C" int64_t sdnq_b200_launch_count(int reset) {
    const int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

// workspace layout: sync [4096 B: strip counters of the fused kernel, zero between launches] | xq [M*K] | sx [M] f32 |
// zx [M] f32 | rowsum [M] i32   (each section 256 B aligned).  The caller zero-fills the first kSyncBytes once, when it
// allocates the workspace; every launch leaves them zero again.
static inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }
static constexpr size_t kSyncBytes = 4096;

extern "C" size_t sdnq_b200_linear_w8a8_workspace_bytes(int64_t M, int64_t K) {
    if (M <= 0 || K <= 0) return 0;
    return kSyncBytes + align256(size_t(M) * size_t(K)) + 3 * align256(size_t(M) * 4);
}

namespace {
struct Workspace {
    int* sync;
    uint8_t* xq;
    float *sx, *zx;
    int32_t* rowsum;
};
int carve_workspace(void* workspace, size_t workspace_bytes, int64_t M, int64_t K, Workspace* w) {
    SDNQ_REQUIRE(workspace != nullptr && workspace_bytes >= sdnq_b200_linear_w8a8_workspace_bytes(M, K), SDNQ_EINVAL,
                 "workspace too small: %zu < %zu", workspace_bytes, sdnq_b200_linear_w8a8_workspace_bytes(M, K));
    SDNQ_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, SDNQ_EINVAL, "workspace must be 256-byte aligned");
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    w->sync = reinterpret_cast<int*>(ws);
    w->xq = ws + kSyncBytes;
    w->sx = reinterpret_cast<float*>(w->xq + align256(size_t(M) * size_t(K)));
    w->zx = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(w->sx) + align256(size_t(M) * 4));
    w->rowsum = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(w->zx) + align256(size_t(M) * 4));
    return SDNQ_OK;
}
// SDNQ_B200_FUSED = 1: linear_w8a8 uses the single-launch kernel wherever it applies; 0 / unset: K2 + K1.  Measured on
// B200 (profiles/r01_fused_vs_split.md) the programmatic-dependent-launch pair is as fast or faster at every SD-XL /
// FLUX shape: the in-kernel producer->consumer hand-off through L2 costs what the kernel boundary costs.
bool fused_by_default() {
    static const bool on = [] {
        const char* e = getenv("SDNQ_B200_FUSED");
        return e && (e[0] == '1' || e[0] == 'y' || e[0] == 'Y' || e[0] == 't' || e[0] == 'T');
    }();
    return on;
}
}  // namespace

extern "C" int sdnq_b200_linear_w8a8_fused(const void* x, int x_dtype, int64_t ldx, const void* wq, int mm_dtype, const float* sw,
                                           const void* bias, int bias_dtype, void* out, int out_dtype, int64_t M, int64_t N, int64_t K,
                                           void* workspace, size_t workspace_bytes, void* stream) {
    SDNQ_REQUIRE(mm_dtype == SDNQ_I8 || mm_dtype == SDNQ_F8E4M3, SDNQ_EUNSUPPORTED, "fused Linear: symmetric int8 / float8_e4m3fn only (got %d)", mm_dtype);
    if (M == 0) return SDNQ_OK;
    Workspace w;
    int rc = carve_workspace(workspace, workspace_bytes, M, K, &w);
    if (rc != SDNQ_OK) return rc;
    rc = linear_fused_impl(x, x_dtype, ldx, wq, mm_dtype, sw, bias, bias_dtype, out, out_dtype, M, N, K, w.xq, w.sx, w.sync,
                           reinterpret_cast<cudaStream_t>(stream));
    SDNQ_REQUIRE(rc <= 0, SDNQ_EUNSUPPORTED,
                 "fused Linear does not cover this case (needs bf16/f16 x == out dtype, K %% 16 == 0, N %% 8 == 0, ldx %% 8 == 0, "
                 "M*K <= 16 Mi, 2*K bytes <= the staging area): M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    return rc;
}

extern "C" int sdnq_b200_linear_w8a8(const void* x, int x_dtype, int64_t ldx, const void* wq, int mm_dtype, const float* sw,
                                     const float* zp, const int32_t* colsum, const void* bias, int bias_dtype, int hadamard_group,
                                     void* out, int out_dtype, int64_t M, int64_t N, int64_t K, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    SDNQ_REQUIRE(mm_dtype == SDNQ_I8 || mm_dtype == SDNQ_U8 || mm_dtype == SDNQ_F8E4M3 || mm_dtype == SDNQ_F8E5M2, SDNQ_EINVAL, "bad matmul dtype %d", mm_dtype);
    const int act_dtype = mm_dtype == SDNQ_F8E5M2 ? SDNQ_F8E4M3 : mm_dtype;      // e5m2 names the weight; activations are always e4m3
    SDNQ_REQUIRE(mm_dtype != SDNQ_U8 || colsum != nullptr, SDNQ_EINVAL, "uint8 matmul needs the weight column sums");
    if (M == 0) return SDNQ_OK;
    Workspace w;
    int rc = carve_workspace(workspace, workspace_bytes, M, K, &w);
    if (rc != SDNQ_OK) return rc;
    const bool need_rowsum = zp != nullptr;
    if (fused_by_default() && hadamard_group == 0 && !need_rowsum && mm_dtype != SDNQ_U8 && mm_dtype != SDNQ_F8E5M2) {
        rc = linear_fused_impl(x, x_dtype, ldx, wq, mm_dtype, sw, bias, bias_dtype, out, out_dtype, M, N, K, w.xq, w.sx, w.sync,
                               reinterpret_cast<cudaStream_t>(stream));
        if (rc <= 0) return rc;
    }
    rc = sdnq_b200_act_quant(x, x_dtype, M, K, ldx, hadamard_group, act_dtype, w.xq, w.sx, mm_dtype == SDNQ_U8 ? w.zx : nullptr,
                             need_rowsum ? w.rowsum : nullptr, nullptr, stream);
    if (rc != SDNQ_OK) return rc;
    return sdnq_b200_scaled_mm(w.xq, wq, (mm_dtype == SDNQ_F8E4M3 || mm_dtype == SDNQ_F8E5M2) ? mm_dtype : SDNQ_I8, w.sx, sw, bias, bias_dtype, 0,
                               need_rowsum ? w.rowsum : nullptr, zp, mm_dtype == SDNQ_U8 ? colsum : nullptr,
                               mm_dtype == SDNQ_U8 ? w.zx : nullptr, out, out_dtype, M, N, K, stream);
}
