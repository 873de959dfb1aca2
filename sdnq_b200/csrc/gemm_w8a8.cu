// K1 host side: tile / pairing choice, argument checks and the C ABI entry points of the W8A8 scaled matmul.  The kernel and its
// launcher template live in gemm_w8a8_kernel.cuh (shared with gemm_w8a8_packed.cu, the instantiations for packed weights other
// than 4-bit integers).
#include "gemm_w8a8_kernel.cuh"

namespace sdnq {
namespace {

// SVD layers: 128-wide single-CTA tiles, generic epilogue, 16-bit outputs (the model dtype of the SVD factors)
template <bool kInt8, int WB>
int launch_gemm_svd(const void* a, const void* b, const GemmParams& p, cudaStream_t st) {
    if (p.out_dtype == SDNQ_BF16) return launch_gemm<128, kInt8, OUT_BF16, false, WB, 0, 1, true>(a, b, p, st);
    return launch_gemm<128, kInt8, OUT_F16, false, WB, 0, 1, true>(a, b, p, st);
}

// packed int4 / uint4 B operand (unpack warps); int8 activations, 128-wide tiles
int launch_gemm_packed4(const void* a, const void* b, const GemmParams& p, cudaStream_t st) {
    if (p.svd_rank != 0) return launch_gemm_svd<true, 4>(a, b, p, st);
    const bool simple = !p.zp && !p.colsum && (p.bias == nullptr || p.bias_ld == 0);
    switch (p.out_dtype) {
        case SDNQ_BF16: return simple ? launch_gemm<128, true, OUT_BF16, true, 4>(a, b, p, st) : launch_gemm<128, true, OUT_BF16, false, 4>(a, b, p, st);
        case SDNQ_F16: return simple ? launch_gemm<128, true, OUT_F16, true, 4>(a, b, p, st) : launch_gemm<128, true, OUT_F16, false, 4>(a, b, p, st);
        default: return simple ? launch_gemm<128, true, OUT_F32, true, 4>(a, b, p, st) : launch_gemm<128, true, OUT_F32, false, 4>(a, b, p, st);
    }
}

// CTA-pair kernels exist for the 2-byte output types (what a bf16 / f16 model produces); everything else runs single-CTA MMAs
template <int BN, bool kInt8>
int launch_gemm_pair(const void* a, const void* b, const GemmParams& p, cudaStream_t st) {
    const bool simple = !p.zp && !p.colsum && (p.bias == nullptr || p.bias_ld == 0);
    if (p.out_dtype == SDNQ_BF16)
        return simple ? launch_gemm<BN, kInt8, OUT_BF16, true, 8, 0, 2>(a, b, p, st) : launch_gemm<BN, kInt8, OUT_BF16, false, 8, 0, 2>(a, b, p, st);
    return simple ? launch_gemm<BN, kInt8, OUT_F16, true, 8, 0, 2>(a, b, p, st) : launch_gemm<BN, kInt8, OUT_F16, false, 8, 0, 2>(a, b, p, st);
}

template <int BN, bool kInt8>
int launch_gemm_out(const void* a, const void* b, const GemmParams& p, cudaStream_t st) {
    if (p.raw) return launch_gemm<BN, kInt8, OUT_RAW32, true>(a, b, p, st);
    const bool simple = !p.zp && !p.colsum && (p.bias == nullptr || p.bias_ld == 0);
    switch (p.out_dtype) {
        case SDNQ_BF16: return simple ? launch_gemm<BN, kInt8, OUT_BF16, true>(a, b, p, st) : launch_gemm<BN, kInt8, OUT_BF16, false>(a, b, p, st);
        case SDNQ_F16: return simple ? launch_gemm<BN, kInt8, OUT_F16, true>(a, b, p, st) : launch_gemm<BN, kInt8, OUT_F16, false>(a, b, p, st);
        default: return simple ? launch_gemm<BN, kInt8, OUT_F32, true>(a, b, p, st) : launch_gemm<BN, kInt8, OUT_F32, false>(a, b, p, st);
    }
}

// Stream-K (gemm_w8a8_kernel<..., kSK>): an int8 GEMM whose 128 x 128 tiles fill the last wave of SMs badly can split its k-blocks
// evenly over all SMs instead.  Needs the caller's workspace (parked partial accumulators + flags).  SDNQ_B200_STREAMK=0 turns it
// off, =1 forces it wherever it is legal (tests, A/B measurements).
size_t stream_k_workspace_bytes() { return size_t(num_sms()) * BM * 128 * 4 + size_t(num_sms()) * 4 * sizeof(int); }

bool use_stream_k(const GemmParams& p, bool i8, int wbits) {
    const char* e = getenv("SDNQ_B200_STREAMK");       // read per call: tests flip it inside one process
    const int mode = e != nullptr ? atoi(e) : -1;
    if (mode == 0 || p.sk_ws == nullptr || !i8 || wbits != 8 || p.raw || p.svd_rank != 0 || p.n_groups != 0) return false;
    if (p.out_dtype != SDNQ_BF16 && p.out_dtype != SDNQ_F16) return false;
    const int sms = num_sms();
    const int64_t tiles = int64_t((p.M + BM - 1) / BM) * ((p.N + 127) / 128), kb = (p.K + BK - 1) / BK;
    if (tiles * kb < 2 * int64_t(sms) || tiles * kb >= (int64_t(1) << 30)) return false;      // at least two k-blocks per CTA
    if (mode == 1) return true;
    const int64_t waves = (tiles + sms - 1) / sms;
    // Measured on B200 (profiles/r02_streamk_ab.log): a CTA streams ~100 GB/s of operands whatever the number of busy SMs (6 stages x
    // 32 KB in flight against ~1.8 us of load latency), and a split tile pays a parked 64 KB accumulator, a flag round trip and an
    // epilogue that nothing overlaps.  Splitting wins only for very deep tiles on a badly filled wave (1024 x 1280 x 10240: 23.6 ->
    // 21.2 us); at SD-XL's K <= 5120 whole tiles are 2 - 5 us faster, so the automatic rule is narrow.
    return tiles * 10 < waves * sms * 8 && tiles <= int64_t(sms) && kb >= 64;
}

template <bool kInt8>
int launch_gemm_sk(const void* a, const void* b, const GemmParams& p, cudaStream_t st) {
    const bool simple = !p.zp && !p.colsum && (p.bias == nullptr || p.bias_ld == 0);
    if (p.out_dtype == SDNQ_BF16)
        return simple ? launch_gemm<128, kInt8, OUT_BF16, true, 8, 0, 1, false, true>(a, b, p, st) : launch_gemm<128, kInt8, OUT_BF16, false, 8, 0, 1, false, true>(a, b, p, st);
    return simple ? launch_gemm<128, kInt8, OUT_F16, true, 8, 0, 1, false, true>(a, b, p, st) : launch_gemm<128, kInt8, OUT_F16, false, 8, 0, 1, false, true>(a, b, p, st);
}

// Tile-N choice: the widest tile that still gives every SM work; BN=256 halves the per-MMA shared-memory
// operand traffic relative to BN=128, so it wins whenever the grid is full either way.
// seg_align != 0 (grouped launch): the segments of the concatenated operand start at multiples of seg_align (128 or 256), which
// limits the tile width to its divisors.
int pick_bn(int M, int N, int seg_align = 0) {
    const char* bn_env = getenv("SDNQ_B200_BN");      // tuning knob, read per call
    const int forced = bn_env ? atoi(bn_env) : 0;
    if (seg_align == 128) return 128;
    if (forced == 128 || (forced == 192 && seg_align == 0) || forced == 256) return forced;
    const int sms = num_sms();
    const int num_m = (M + BM - 1) / BM;
    auto waves_eff = [&](int bn) {
        const int tiles = num_m * ((N + bn - 1) / bn);
        const int waves = (tiles + sms - 1) / sms;
        return static_cast<double>(tiles) / (static_cast<double>(waves) * sms);
    };
    // relative per-tile efficiency of the narrower tiles (shared-memory operand traffic per MMA grows as BN shrinks)
    const double e256 = waves_eff(256) * 1.00, e192 = seg_align != 0 ? 0.0 : waves_eff(192) * 0.94, e128 = waves_eff(128) * 0.90;
    if (e256 >= e192 && e256 >= e128) return 256;
    return e192 >= e128 ? 192 : 128;
}

// CTA pairs (tcgen05 cta_group::2): 0 = use single-CTA MMAs, else the tile width.  SDNQ_B200_CG=1 / 2 forces the choice.
int pick_pair(const GemmParams& p) {
    const char* cg_env = getenv("SDNQ_B200_CG");      // read per call: tests and A/B tools flip it inside one process
    const int forced = cg_env ? atoi(cg_env) : 0;
    if (forced == 1 || p.raw || (p.out_dtype != SDNQ_BF16 && p.out_dtype != SDNQ_F16) || p.M <= BM) return 0;
    for (int g = 1; g <= p.n_groups; ++g)
        if (p.grp_start[g] % 256 != 0) return 0;                          // grouped segments on a 128 grid: 128-wide single-CTA tiles
    const char* bn_env = getenv("SDNQ_B200_BN");
    const int forced_bn = bn_env ? atoi(bn_env) : 0;
    const int pairs = num_sms() / 2;
    const int num_m2 = (p.M + 2 * BM - 1) / (2 * BM);
    auto eff = [&](int bn) {
        const int tiles = num_m2 * ((p.N + bn - 1) / bn);
        const int waves = (tiles + pairs - 1) / pairs;
        return static_cast<double>(tiles) / (static_cast<double>(waves) * pairs);
    };
    if (forced == 2) return forced_bn == 128 || forced_bn == 256 ? forced_bn : (eff(256) >= 0.92 * eff(128) ? 256 : 128);
    // Auto: a small cost model calibrated on B200 (profiles/r01_gemm_pair_vs_single.md; microseconds, k-block = 128 bytes of K):
    //   time = waves * (k-blocks * per-k-block cost + exposed epilogue) + launch / prologue / drain
    // single-CTA tiles stream 16 KB + BN * 128 B of operands per k-block and SM, a pair 16 KB + 128 * 128 B for a 256-wide tile,
    // which is what makes the pair faster once a launch has several waves of deep tiles and slower for the small SD-XL GEMMs
    // (cluster launch + two cluster barriers, half as many schedulable units).
    const int kb = (p.K + BK - 1) / BK;
    const int sms = num_sms();
    const int num_m = (p.M + BM - 1) / BM;
    auto single_time = [&](int bn, double per_kb) {
        const int tiles = num_m * ((p.N + bn - 1) / bn);
        return ((tiles + sms - 1) / sms) * (kb * per_kb + 0.3) + 4.5;
    };
    const double t_single = fmin(single_time(256, 0.40), fmin(single_time(192, 0.32), single_time(128, 0.25)));
    const int ptiles = num_m2 * ((p.N + 255) / 256);
    const double t_pair = ((ptiles + pairs - 1) / pairs) * (kb * 0.36 + 0.3) + 6.0;
    return t_pair < 0.97 * t_single ? 256 : 0;
}

}  // namespace

int scaled_mm_impl(const void* a, const void* b, int ab_dtype, GemmParams p, cudaStream_t st, int wbits = 8) {
    SDNQ_REQUIRE(a && b && p.out, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(ab_dtype == SDNQ_I8 || ab_dtype == SDNQ_F8E4M3 || ab_dtype == SDNQ_F8E5M2, SDNQ_EINVAL,
                 "operand dtype must be int8, float8_e4m3fn or float8_e5m2 (= e4m3 activations x e5m2 weights) (got %d)", ab_dtype);
    SDNQ_REQUIRE(p.M >= 0 && p.N > 0 && p.K > 0, SDNQ_EINVAL, "bad shape M=%d N=%d K=%d", p.M, p.N, p.K);
    SDNQ_REQUIRE(p.K % 16 == 0, SDNQ_EUNSUPPORTED, "K (=%d) must be a multiple of 16 (TMA row pitch)", p.K);
    SDNQ_REQUIRE(p.N % 8 == 0, SDNQ_EUNSUPPORTED, "N (=%d) must be a multiple of 8 (16-byte output vectors)", p.N);
    SDNQ_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(p.out) & 15) == 0, SDNQ_EINVAL, "a, b and out must be 16-byte aligned");
    if (!p.raw) {
        SDNQ_REQUIRE(p.sx && p.sw, SDNQ_EINVAL, "scale pointer is NULL");
        SDNQ_REQUIRE(p.out_dtype == SDNQ_BF16 || p.out_dtype == SDNQ_F16 || p.out_dtype == SDNQ_F32, SDNQ_EINVAL, "bad out dtype %d", p.out_dtype);
        if (p.bias) SDNQ_REQUIRE(p.bias_dtype == SDNQ_BF16 || p.bias_dtype == SDNQ_F16 || p.bias_dtype == SDNQ_F32, SDNQ_EINVAL, "bad bias dtype %d", p.bias_dtype);
        SDNQ_REQUIRE(!p.zp || p.rowsum || p.colsum, SDNQ_EINVAL, "zp given without rowsum");
        SDNQ_REQUIRE(!p.colsum || p.zx, SDNQ_EINVAL, "colsum given without zx");
    }
    if (p.svd_rank != 0) {
        SDNQ_REQUIRE(!p.raw && (p.out_dtype == SDNQ_BF16 || p.out_dtype == SDNQ_F16), SDNQ_EUNSUPPORTED, "the SVD accumulate needs a bf16 / f16 output");
        SDNQ_REQUIRE(p.svd_rank == 16 || p.svd_rank == 32 || p.svd_rank == 64, SDNQ_EUNSUPPORTED, "svd rank must be 16, 32 or 64 (got %d)", p.svd_rank);
        SDNQ_REQUIRE(p.svd_low && p.svd_up && ((reinterpret_cast<uintptr_t>(p.svd_low) | reinterpret_cast<uintptr_t>(p.svd_up)) & 15) == 0, SDNQ_EINVAL,
                     "svd operands must be non-NULL and 16-byte aligned");
    }
    if (p.M == 0) return SDNQ_OK;
    const bool i8 = ab_dtype == SDNQ_I8;
    if (wbits != 8) {
        SDNQ_REQUIRE(wbits >= 2 && wbits <= 7 && !p.raw, SDNQ_EUNSUPPORTED, "in-kernel unpack covers 2..7-bit weights (got %d bits)", wbits);
        SDNQ_REQUIRE((int64_t(p.K) * wbits) % 128 == 0, SDNQ_EUNSUPPORTED, "packed %d-bit B needs K * bits %% 128 == 0 (16-byte row pitch), K=%d", wbits, p.K);
        if (wbits == 4 && p.pk_kind == 0) {
            SDNQ_REQUIRE(i8, SDNQ_EUNSUPPORTED, "packed integer weights need int8 activations");
            return launch_gemm_packed4(a, b, p, st);
        }
        SDNQ_REQUIRE(p.svd_rank == 0, SDNQ_EUNSUPPORTED, "the SVD accumulate takes 4-bit or unpacked weights");
        SDNQ_REQUIRE(p.pk_kind == 0 ? i8 : ab_dtype == SDNQ_F8E4M3, SDNQ_EUNSUPPORTED, "packed integer weights need int8 activations, packed minifloats float8_e4m3fn ones");
        return launch_gemm_packed_any(a, b, p, st);
    }
    if (p.svd_rank != 0) return i8 ? launch_gemm_svd<true, 8>(a, b, p, st) : launch_gemm_svd<false, 8>(a, b, p, st);
    if (use_stream_k(p, i8, wbits)) return launch_gemm_sk<true>(a, b, p, st);
    if (const int pair_bn = pick_pair(p); pair_bn != 0) {
        if (pair_bn == 256) return i8 ? launch_gemm_pair<256, true>(a, b, p, st) : launch_gemm_pair<256, false>(a, b, p, st);
        return i8 ? launch_gemm_pair<128, true>(a, b, p, st) : launch_gemm_pair<128, false>(a, b, p, st);
    }
    int seg_align = 0;
    if (p.n_groups != 0) {
        seg_align = 256;
        for (int g = 0; g <= p.n_groups; ++g)
            if (p.grp_start[g] % 256 != 0) seg_align = 128;
    }
    switch (pick_bn(p.M, p.N, seg_align)) {
        case 256: return i8 ? launch_gemm_out<256, true>(a, b, p, st) : launch_gemm_out<256, false>(a, b, p, st);
        case 192: return i8 ? launch_gemm_out<192, true>(a, b, p, st) : launch_gemm_out<192, false>(a, b, p, st);
        case 128: return i8 ? launch_gemm_out<128, true>(a, b, p, st) : launch_gemm_out<128, false>(a, b, p, st);
        default: return i8 ? launch_gemm_out<128, true>(a, b, p, st) : launch_gemm_out<128, false>(a, b, p, st);
    }
}

// Fused W8A8 Linear (one launch): quantise x [M,K] (bf16 / f16) per row into xq / sx and run the scaled GEMM on it.
// Returns 1 when the problem is outside what the fused kernel covers (the caller then runs K2 + K1).
//   covered: symmetric int8 / e4m3 activations without Hadamard, no zero-point terms, vector (or no) bias,
//            out dtype == x dtype, one row of x fits the staging area, <= kSyncStrips row strips.
int linear_fused_impl(const void* x, int x_dtype, int64_t ldx, const void* wq, int ab_dtype, const float* sw, const void* bias,
                      int bias_dtype, void* out, int out_dtype, int64_t M, int64_t N, int64_t K, uint8_t* xq, float* sx, int* sync,
                      cudaStream_t st) {
    if (!(x_dtype == SDNQ_BF16 || x_dtype == SDNQ_F16) || out_dtype != x_dtype) return 1;
    if (ab_dtype != SDNQ_I8 && ab_dtype != SDNQ_F8E4M3 && ab_dtype != SDNQ_F8E5M2) return 1;
    if (M <= 0 || M > int64_t(kSyncStrips) * BM || K % 16 != 0 || N % 8 != 0 || ldx % 8 != 0) return 1;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return 1;
    // per-CTA share of x beyond a few staging batches: the quantiser is no longer latency-bound and the
    // stand-alone pre-pass (more warps, registers instead of shared memory) is the better tool
    if (M * K > (int64_t(16) << 20)) return 1;
    const int bn = pick_bn(int(M), int(N));
    const int64_t stage_bytes = bn == 256 ? int64_t(Cfg<256>::kStages) * Cfg<256>::kStageA : int64_t(Cfg<128>::kStages) * Cfg<128>::kStageA;
    if (K * 2 > stage_bytes) return 1;
    GemmParams p{sx, sw, bias, bias_dtype, 0, nullptr, nullptr, nullptr, nullptr, out, out_dtype, int(M), int(N), int(K), 0, 0u, ab_dtype == SDNQ_F8E5M2 ? 1u : 0u, nullptr, nullptr, 0, 0u, x, ldx, xq, sx, sync};
    const bool i8 = ab_dtype == SDNQ_I8;
    const bool bf = x_dtype == SDNQ_BF16;
#define SDNQ_FUSED(BN_)                                                                                                    \
    (i8 ? (bf ? launch_gemm<BN_, true, OUT_BF16, true, 8, 1>(xq, wq, p, st) : launch_gemm<BN_, true, OUT_F16, true, 8, 2>(xq, wq, p, st))  \
        : (bf ? launch_gemm<BN_, false, OUT_BF16, true, 8, 1>(xq, wq, p, st) : launch_gemm<BN_, false, OUT_F16, true, 8, 2>(xq, wq, p, st)))
    return bn == 256 ? SDNQ_FUSED(256) : SDNQ_FUSED(128);
#undef SDNQ_FUSED
}

}  // namespace sdnq

using namespace sdnq;

extern "C" int sdnq_b200_scaled_mm(const void* a, const void* b, int ab_dtype, const float* sx, const float* sw, const void* bias,
                                   int bias_dtype, int64_t bias_ld, const int32_t* rowsum, const float* zp, const int32_t* colsum,
                                   const float* zx, void* out, int out_dtype, int64_t M, int64_t N, int64_t K, void* stream) {
    SDNQ_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), SDNQ_EUNSUPPORTED, "dimension too large");
    GemmParams p{sx, sw, bias, bias_dtype, bias_ld, rowsum, zp, colsum, zx, out, out_dtype, (int)M, (int)N, (int)K, 0, 0u, ab_dtype == SDNQ_F8E5M2 ? 1u : 0u};
    return scaled_mm_impl(a, b, ab_dtype, p, reinterpret_cast<cudaStream_t>(stream));
}

// packed storage format -> the GEMM's prologue parameters; *ab = operand dtype of the activations, *wbits = storage bits
static int packed_params(const sdnq_weight_format* b_fmt, const float* zp, const int32_t* rowsum, GemmParams& p, int* ab, int* wbits) {
    WFormat f;
    int rc = make_wformat(b_fmt, &f);
    if (rc != SDNQ_OK) return rc;
    SDNQ_REQUIRE(f.bits >= 2 && f.bits <= 7 && f.word_bytes == 1 && (f.kind == SDNQ_W_INT || f.kind == SDNQ_W_MINIFLOAT), SDNQ_EUNSUPPORTED,
                 "packed weights: 2..7-bit integer or minifloat codes are expanded in-kernel (got kind %d, %d bits)", f.kind, f.bits);
    p.pk_bits = *wbits = f.bits;
    if (f.kind == SDNQ_W_INT) {
        SDNQ_REQUIRE(f.is_unsigned == 0 || (zp != nullptr && rowsum != nullptr), SDNQ_EINVAL, "unsigned integer weights need zp and rowsum");
        p.pk_kind = 0;
        p.w_sub = f.is_unsigned ? 0u : 0x01010101u * static_cast<uint32_t>(1 << (f.bits - 1));     // offset-binary -> two's complement
        *ab = SDNQ_I8;
        return SDNQ_OK;
    }
    SDNQ_REQUIRE(f.exponent <= 4 && f.mantissa <= 3, SDNQ_EUNSUPPORTED, "packed weights: e%dm%d is not a subset of e4m3", f.exponent, f.mantissa);
    p.pk_kind = 1;
    p.pk_exp = f.exponent;
    p.pk_man = f.mantissa;
    p.pk_unsigned = f.is_unsigned;
    *ab = SDNQ_F8E4M3;
    return SDNQ_OK;
}

extern "C" size_t sdnq_b200_scaled_mm_workspace_bytes(void) { return stream_k_workspace_bytes(); }

extern "C" int sdnq_b200_scaled_mm_ws(const void* a, const void* b, int ab_dtype, const float* sx, const float* sw, const void* bias,
                                      int bias_dtype, int64_t bias_ld, const int32_t* rowsum, const float* zp, const int32_t* colsum,
                                      const float* zx, void* out, int out_dtype, int64_t M, int64_t N, int64_t K, void* workspace,
                                      size_t workspace_bytes, void* stream) {
    SDNQ_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), SDNQ_EUNSUPPORTED, "dimension too large");
    GemmParams p{sx, sw, bias, bias_dtype, bias_ld, rowsum, zp, colsum, zx, out, out_dtype, (int)M, (int)N, (int)K, 0, 0u, ab_dtype == SDNQ_F8E5M2 ? 1u : 0u};
    if (workspace != nullptr && workspace_bytes >= stream_k_workspace_bytes() && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0) {
        p.sk_flags = reinterpret_cast<int*>(workspace);                                              // [SMs][4], zero between launches
        p.sk_ws = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(workspace) + size_t(num_sms()) * 4 * sizeof(int));
    }
    return scaled_mm_impl(a, b, ab_dtype, p, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int sdnq_b200_scaled_mm_packed(const void* a, const void* b_packed, const sdnq_weight_format* b_fmt, const float* sx, const float* sw,
                                          const void* bias, int bias_dtype, int64_t bias_ld, const int32_t* rowsum, const float* zp, void* out,
                                          int out_dtype, int64_t M, int64_t N, int64_t K, void* stream) {
    SDNQ_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), SDNQ_EUNSUPPORTED, "dimension too large");
    GemmParams p{sx, sw, bias, bias_dtype, bias_ld, rowsum, zp, nullptr, nullptr, out, out_dtype, (int)M, (int)N, (int)K, 0, 0u, 0u};
    int ab = SDNQ_I8, wbits = 8;
    int rc = packed_params(b_fmt, zp, rowsum, p, &ab, &wbits);
    if (rc != SDNQ_OK) return rc;
    return scaled_mm_impl(a, b_packed, ab, p, reinterpret_cast<cudaStream_t>(stream), wbits);
}

extern "C" int sdnq_b200_mm(const void* a, const void* b, int ab_dtype, void* out, int64_t M, int64_t N, int64_t K, void* stream) {
    SDNQ_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), SDNQ_EUNSUPPORTED, "dimension too large");
    GemmParams p{nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, out, SDNQ_I32, (int)M, (int)N, (int)K, 1, 0u, ab_dtype == SDNQ_F8E5M2 ? 1u : 0u};
    return scaled_mm_impl(a, b, ab_dtype, p, reinterpret_cast<cudaStream_t>(stream));
}

// ---- scaled matmul with the SVD rank-r term accumulated on the tensor cores (the SVD branch of get_*_matmul_inputs)
extern "C" int sdnq_b200_scaled_mm_svd(const void* a, const void* b, int ab_dtype, const sdnq_weight_format* b_fmt, const float* sx, const float* sw,
                                       const void* bias, int bias_dtype, int64_t bias_ld, const int32_t* rowsum, const float* zp,
                                       const int32_t* colsum, const float* zx, const void* svd_low, const void* svd_up_nr, int svd_rank,
                                       int svd_dtype, void* out, int out_dtype, int64_t M, int64_t N, int64_t K, void* stream) {
    SDNQ_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), SDNQ_EUNSUPPORTED, "dimension too large");
    SDNQ_REQUIRE(svd_dtype == SDNQ_BF16 || svd_dtype == SDNQ_F16, SDNQ_EUNSUPPORTED, "svd factors must be bf16 / f16 (got %d)", svd_dtype);
    SDNQ_REQUIRE(svd_rank != 0, SDNQ_EINVAL, "svd_rank is 0: use sdnq_b200_scaled_mm");
    uint32_t w_sub = 0u;
    int wbits = 8;
    if (b_fmt != nullptr) {
        SDNQ_REQUIRE(b_fmt->kind == SDNQ_W_INT && b_fmt->bits == 4, SDNQ_EUNSUPPORTED, "scaled_mm_svd: packed weights must be int4 / uint4");
        SDNQ_REQUIRE(b_fmt->is_unsigned == 0 || (zp != nullptr && rowsum != nullptr), SDNQ_EINVAL, "uint4 weights need zp and rowsum");
        w_sub = b_fmt->is_unsigned ? 0u : 0x08080808u;
        wbits = 4;
    }
    GemmParams p{sx, sw, bias, bias_dtype, bias_ld, rowsum, zp, colsum, zx, out, out_dtype, (int)M, (int)N, (int)K, 0, w_sub,
                 ab_dtype == SDNQ_F8E5M2 ? 1u : 0u, svd_low, svd_up_nr, svd_rank, svd_dtype == SDNQ_BF16 ? 1u : 0u};
    return scaled_mm_impl(a, b, wbits == 4 ? SDNQ_I8 : ab_dtype, p, reinterpret_cast<cudaStream_t>(stream), wbits);
}

// ---- grouped scaled matmul: sibling projections (to_q / to_k / to_v ...) that read the same quantised activations, one launch
extern "C" int sdnq_b200_scaled_mm_grouped(const void* a, const void* b_cat, int ab_dtype, const sdnq_weight_format* b_fmt, const float* sx,
                                           const float* sw_cat, const void* bias_cat, int bias_dtype, const int32_t* rowsum, const float* zp_cat,
                                           const int32_t* colsum_cat, const float* zx, int n_groups, const int64_t* seg_start, const int64_t* seg_n,
                                           void* const* outs, int out_dtype, int64_t M, int64_t K, void* stream) {
    SDNQ_REQUIRE(n_groups >= 1 && n_groups <= kMaxGroups, SDNQ_EUNSUPPORTED, "grouped launch: 1..%d siblings (got %d)", kMaxGroups, n_groups);
    SDNQ_REQUIRE(seg_start && seg_n && outs, SDNQ_EINVAL, "NULL pointer");
    SDNQ_REQUIRE(M < (1LL << 31) && K < (1LL << 31) && seg_start[n_groups] < (1LL << 31), SDNQ_EUNSUPPORTED, "dimension too large");
    GemmParams p{sx, sw_cat, bias_cat, bias_dtype, 0, rowsum, zp_cat, colsum_cat, zx, outs[0], out_dtype, (int)M, (int)seg_start[n_groups], (int)K, 0, 0u,
                 ab_dtype == SDNQ_F8E5M2 ? 1u : 0u};
    int wbits = 8;
    if (b_fmt != nullptr) {
        int rc = packed_params(b_fmt, zp_cat, rowsum, p, &ab_dtype, &wbits);
        if (rc != SDNQ_OK) return rc;
    }
    p.n_groups = n_groups;
    SDNQ_REQUIRE(seg_start[0] == 0, SDNQ_EINVAL, "grouped launch: the first segment starts at 0");
    for (int g = 0; g < n_groups; ++g) {
        SDNQ_REQUIRE(seg_n[g] > 0 && seg_n[g] % 8 == 0 && seg_start[g] % 128 == 0 && seg_start[g] + seg_n[g] <= seg_start[g + 1], SDNQ_EINVAL,
                     "grouped launch: segment %d (start %lld, %lld rows) must start at a multiple of 128, hold a multiple of 8 rows and end before the next one",
                     g, (long long)seg_start[g], (long long)seg_n[g]);
        SDNQ_REQUIRE(outs[g] && (reinterpret_cast<uintptr_t>(outs[g]) & 15) == 0, SDNQ_EINVAL, "grouped launch: output %d is NULL or not 16-byte aligned", g);
        p.grp_start[g] = (int)seg_start[g];
        p.grp_n[g] = (int)seg_n[g];
        p.grp_out[g] = outs[g];
    }
    SDNQ_REQUIRE(seg_start[n_groups] % 128 == 0, SDNQ_EINVAL, "grouped launch: the concatenated operand must end at a multiple of 128 rows");
    p.grp_start[n_groups] = (int)seg_start[n_groups];
    return scaled_mm_impl(a, b_cat, ab_dtype, p, reinterpret_cast<cudaStream_t>(stream), wbits);
}
