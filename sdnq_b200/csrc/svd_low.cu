// K7: the rank-r projection of the SVD branch of a W8A8 forward,  low[M,r] = cast_T( x[M,K] @ svd_down[K,r] ).
//
// Reference behaviour restated here: get_int8_matmul_inputs / get_uint8_matmul_inputs / get_fp8_matmul_inputs add the SVD term
// as a dense bias computed on the rotated, un-quantised activations in the SVD dtype (layers/linear/linear_int8.py:57-62):
//     bias2d = bias + torch.mm(torch.mm(x, svd_down), svd_up)
// i.e. two library GEMMs and an [M,N] tensor written and re-read per call.  Here the first product is this kernel (rounded to the
// activation dtype exactly where torch.mm rounds it) and the second one is a rank-r tcgen05 accumulate inside the scaled GEMM
// (gemm_w8a8.cu), so neither the [M,N] bias nor a library GEMM remains on the path.
//
// Mapping: CTA = 16 rows of x; its 8 warps split K in 64-column steps and reduce through shared memory.  mma.sync.m16n8k16
// (f32 accumulate) with the fragment k-slots permuted so that lane (g, t) owns the 16 consecutive columns [16t, 16t+16) of a
// step for both operands: two 16-byte loads per x row, two per svd_down row, four MMAs per 8 output columns.  svd_down is
// passed as [r, K] row-major (a cached transpose of the stored [K, r] factor: r*K*2 bytes).  Memory-bound on x (L2-resident: the
// activation quantiser has just written it); algorithmic bytes 2*M*K + 2*r*K + 2*M*r.
#include "hadamard_tc.cuh"     // hadtc::Half16<T>: cvt pack + mma.sync m16n8k16 (with their host-emulation models)

namespace sdnq {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxNT = 8;          // rank <= 64

struct LowArgs {
    const void* x;                 // [M, K] of T, row stride ldx
    int64_t ldx;
    const void* down;              // [rank, K] of T, row-major
    void* low;                     // [M, rank] of T
    int M, K, rank;
};

template <typename T>
__global__ void __launch_bounds__(kThreads) svd_low_kernel(const LowArgs a) {
    __shared__ float s_red[kWarps - 1][kMaxNT * 4][32];
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nt = a.rank / 8;
    const T* x = reinterpret_cast<const T*>(a.x);
    const T* down = reinterpret_cast<const T*>(a.down);
    const int row_lo = blockIdx.x * 16 + g, row_hi = row_lo + 8;
    const bool lo_ok = row_lo < a.M, hi_ok = row_hi < a.M;
    const int steps = (a.K + 63) / 64;
    float acc[kMaxNT][4];
#pragma unroll
    for (int j = 0; j < kMaxNT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    // one step = 64 columns of K: per lane 2 x 16 B of each of its two x rows and 2 x 16 B of each svd_down row it feeds.  The loads of
    // step s + kWarps are issued before the MMAs of step s (register double buffer): a warp walks K / 512 dependent steps, and
    // without the prefetch every one of them exposes a full L2 round trip.
    struct Frag { uint4 xl0, xl1, xh0, xh1, d0[kMaxNT], d1[kMaxNT]; };
    auto load = [&](int s, Frag& f) {
        const int k = s * 64 + 16 * t;
        const bool live = s < steps && k < a.K;          // K % 16 == 0: a lane's 16 columns are all inside or all outside
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        f.xl0 = f.xl1 = f.xh0 = f.xh1 = z;
        if (live && lo_ok) {
            const T* p = x + int64_t(row_lo) * a.ldx + k;
            f.xl0 = *reinterpret_cast<const uint4*>(p);
            f.xl1 = *reinterpret_cast<const uint4*>(p + 8);
        }
        if (live && hi_ok) {
            const T* p = x + int64_t(row_hi) * a.ldx + k;
            f.xh0 = *reinterpret_cast<const uint4*>(p);
            f.xh1 = *reinterpret_cast<const uint4*>(p + 8);
        }
#pragma unroll
        for (int j = 0; j < kMaxNT; ++j) {
            f.d0[j] = f.d1[j] = z;
            if (j < nt && live) {
                const T* p = down + int64_t(8 * j + g) * a.K + k;      // B fragment: column n = g of this 8-column block
                f.d0[j] = *reinterpret_cast<const uint4*>(p);
                f.d1[j] = *reinterpret_cast<const uint4*>(p + 8);
            }
        }
    };
    auto contract = [&](const Frag& f) {
#pragma unroll
        for (int j = 0; j < kMaxNT; ++j) {
            if (j < nt) {
                // MMA i contracts the lane's elements [4i, 4i+4): slots (2t, 2t+1) <- elements 4i, 4i+1; (2t+8, 2t+9) <- 4i+2, 4i+3
                hadtc::Half16<T>::mma(acc[j], f.xl0.x, f.xh0.x, f.xl0.y, f.xh0.y, f.d0[j].x, f.d0[j].y);
                hadtc::Half16<T>::mma(acc[j], f.xl0.z, f.xh0.z, f.xl0.w, f.xh0.w, f.d0[j].z, f.d0[j].w);
                hadtc::Half16<T>::mma(acc[j], f.xl1.x, f.xh1.x, f.xl1.y, f.xh1.y, f.d1[j].x, f.d1[j].y);
                hadtc::Half16<T>::mma(acc[j], f.xl1.z, f.xh1.z, f.xl1.w, f.xh1.w, f.d1[j].z, f.d1[j].w);
            }
        }
    };
    Frag fa, fb;
    load(warp, fa);
    for (int s = warp; s < steps; s += 2 * kWarps) {
        load(s + kWarps, fb);
        contract(fa);
        if (s + kWarps >= steps) break;
        load(s + 2 * kWarps, fa);
        contract(fb);
    }
    // ---- add the K-split partials up in warp 0
    if (warp > 0) {
#pragma unroll
        for (int j = 0; j < kMaxNT; ++j)
            if (j < nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) s_red[warp - 1][j * 4 + i][lane] = acc[j][i];
    }
    __syncthreads();
    if (warp != 0) return;
#pragma unroll
    for (int j = 0; j < kMaxNT; ++j) {
        if (j < nt) {
#pragma unroll
            for (int w = 0; w < kWarps - 1; ++w)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[j][i] += s_red[w][j * 4 + i][lane];
            // C fragment: (row g | g+8) x (columns 8j + 2t, 8j + 2t + 1)
            uint32_t* out = reinterpret_cast<uint32_t*>(a.low);
            if (lo_ok) out[(int64_t(row_lo) * a.rank + 8 * j + 2 * t) >> 1] = hadtc::Half16<T>::pack(acc[j][0], acc[j][1]);
            if (hi_ok) out[(int64_t(row_hi) * a.rank + 8 * j + 2 * t) >> 1] = hadtc::Half16<T>::pack(acc[j][2], acc[j][3]);
        }
    }
}

}  // namespace
}  // namespace sdnq

using namespace sdnq;

extern "C" int sdnq_b200_svd_low(const void* x, int x_dtype, int64_t ldx, const void* svd_down_rk, int svd_rank, void* low,
                                 int64_t M, int64_t K, void* stream) {
    SDNQ_REQUIRE(x && svd_down_rk && low, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(x_dtype == SDNQ_BF16 || x_dtype == SDNQ_F16, SDNQ_EUNSUPPORTED, "svd_low: bf16 / f16 activations (got %d)", x_dtype);
    SDNQ_REQUIRE(M >= 0 && K > 0 && M < (1LL << 31) && K < (1LL << 31), SDNQ_EINVAL, "bad shape M=%lld K=%lld", (long long)M, (long long)K);
    SDNQ_REQUIRE(svd_rank >= 8 && svd_rank <= 8 * kMaxNT && svd_rank % 8 == 0, SDNQ_EUNSUPPORTED, "svd_low: rank must be a multiple of 8 up to 64 (got %d)", svd_rank);
    SDNQ_REQUIRE(K % 16 == 0 && ldx % 8 == 0 && ldx >= K, SDNQ_EUNSUPPORTED, "svd_low: K %% 16 == 0 and ldx %% 8 == 0 (K=%lld ldx=%lld)", (long long)K, (long long)ldx);
    SDNQ_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(svd_down_rk)) & 15) == 0 && (reinterpret_cast<uintptr_t>(low) & 3) == 0, SDNQ_EINVAL,
                 "x and svd_down must be 16-byte aligned");
    if (M == 0) return SDNQ_OK;
    LowArgs a{x, ldx, svd_down_rk, low, int(M), int(K), svd_rank};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = int((M + 15) / 16);
    cudaError_t e = x_dtype == SDNQ_BF16 ? launch_pdl(svd_low_kernel<__nv_bfloat16>, dim3(grid), dim3(kThreads), 0, st, a)
                                         : launch_pdl(svd_low_kernel<__half>, dim3(grid), dim3(kThreads), 0, st, a);
    if (e != cudaSuccess) return set_error(SDNQ_ECUDA, "launch of svd_low_kernel failed: %s", cudaGetErrorString(e));
    return check_launch("svd_low_kernel");
}
