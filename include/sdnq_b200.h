/*
 * sdnq_b200.h -- C ABI of the B200-native SDNQ quantized-Linear / Conv kernels.
 *
 * Every entry point is what a binding of the reference's kernel layer for this
 * path would call; the reference interface each one replaces is cited as
 * file:line under /root/reference/src/sdnq/.  Plain pointers and sizes only (no
 * torch / C++ types).  All pointers are DEVICE pointers unless the name ends in
 * `_host`; the caller owns every buffer including outputs and workspaces; no
 * entry point allocates, frees, or synchronises; all of them are asynchronous on
 * `stream` (a `cudaStream_t` passed as `void*`) and CUDA-graph capturable.
 *
 * Stream ordering: kernels are launched with programmatic dependent launch; every kernel waits (griddepcontrol.wait) before it
 * reads anything its stream predecessors may have produced -- with one exception: the GEMM entry points prefetch *weight*
 * tiles (b / wq) before that wait.  Weights are frozen in this path; the library orders its own operand producers
 * (sdnq_b200_unpack, sdnq_b200_requant) with a trailing fence kernel, and a caller that rewrites a weight buffer must let that
 * write complete (any ordinary kernel or copy in between suffices) before the next GEMM call on the stream.
 *
 * Return value: 0 on success, negative `sdnq_status` otherwise; the message of
 * the last failure on the calling thread is returned by `sdnq_b200_last_error()`.
 * There is no CPU fallback anywhere behind this ABI.
 */
#ifndef SDNQ_B200_H
#define SDNQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDNQ_B200_ABI_VERSION 2

#if defined(__GNUC__)
#define SDNQ_API __attribute__((visibility("default")))
#else
#define SDNQ_API
#endif

typedef enum sdnq_status {
    SDNQ_OK = 0,
    SDNQ_EINVAL = -1,       /* bad argument (shape / alignment / enum) */
    SDNQ_EUNSUPPORTED = -2, /* valid request the kernels do not cover (message says which) */
    SDNQ_ECUDA = -3,        /* a CUDA runtime / driver call failed */
    SDNQ_EARCH = -4         /* device is not sm_100 */
} sdnq_status;

/* element types of activations / outputs / scales */
typedef enum sdnq_dtype {
    SDNQ_F32 = 0,
    SDNQ_BF16 = 1,
    SDNQ_F16 = 2,
    SDNQ_I8 = 3,
    SDNQ_U8 = 4,
    SDNQ_F8E4M3 = 5,
    SDNQ_I32 = 6,
    SDNQ_F8E5M2 = 7 /* weight-side only: as the operand dtype of a matmul entry point it names the mixed pair
                       a = float8_e4m3fn activations x b = float8_e5m2 weights (what the reference hands to
                       torch._scaled_mm for weights_dtype="float8_e5m2", linear_fp8.py:25-54) */
} sdnq_dtype;

/* storage format of a quantised weight: one row of reference common.py:16-267 (`dtype_dict`) */
typedef enum sdnq_wkind {
    SDNQ_W_INT = 0,       /* intN / uintN, N = 1..8; N < 8 packed as packed_int/pack.py lays it out;
                             signed types are stored offset-binary (packed_int/__init__.py:76-88) */
    SDNQ_W_MINIFLOAT = 1, /* floatB_eXmY "fn"/"fnu", B = 1..8, packed like uintB (packed_float.py:26-132) */
    SDNQ_W_FP8_E4M3FN = 2, /* torch.float8_e4m3fn, unpacked */
    SDNQ_W_FP8_E5M2 = 3    /* torch.float8_e5m2, unpacked */
} sdnq_wkind;

typedef struct sdnq_weight_format {
    int32_t kind;        /* sdnq_wkind */
    int32_t bits;        /* 1..8 */
    int32_t is_unsigned; /* uintN / "fnu" */
    int32_t exponent;    /* minifloat exponent bits */
    int32_t mantissa;    /* minifloat mantissa bits */
    int32_t word_bytes;  /* bytes per storage word: 1, or 8 for uint1 (the reference stores one int64 per
                            packed byte because bool shifts promote, packed_int/pack.py:309-321) */
} sdnq_weight_format;

/* ---- library ------------------------------------------------------------------------------- */
SDNQ_API int sdnq_b200_abi_version(void);
SDNQ_API const char* sdnq_b200_last_error(void);
/* 0 if device `device` can run these kernels (compute capability 10.x), SDNQ_EARCH otherwise */
SDNQ_API int sdnq_b200_check_device(int device);

/* ---- unpack:  packed_int.unpack_int (packed_int/__init__.py:83-88) and
 *               packed_float.unpack_float (packed_float.py:85-132) ------------------------------
 * packed storage -> `numel` codes / values in flattened row-major order.
 * out_dtype: SDNQ_I8 / SDNQ_U8 / SDNQ_I32 / SDNQ_F32 for integer formats (signed offset removed),
 *            SDNQ_F32 / SDNQ_BF16 / SDNQ_F8E4M3 for minifloat formats.  numel % 8 == 0. */
SDNQ_API int sdnq_b200_unpack(const void* packed, const sdnq_weight_format* fmt, void* out, int out_dtype,
                     int64_t numel, void* stream);

/* ---- K3 dequant:  SDNQDequantizer.__call__ / dequantize_weight (dequantizer.py:135-162, 389-429)
 * W[N,K] = cast( unpack(weight) * scale [+ zero_point] ) [+ svd_up @ svd_down] [un-rotate]
 *   weight      physical [N,K]-ordered storage (packed or not; a K-major [K,N] matmul-layout tensor
 *               of the reference is physically this, quant_utils.py:239-249)
 *   scale/zp    f32, one per (row, K-group): physical [N, K/group]; group_size = K for row-wise,
 *               group_size = -2 tensor-wise (a single scalar).  zp NULL for symmetric formats.
 *   codebook    if non-zero, `scale` holds levels [N, K/group, 2^bits] and codes index it
 *               (dequantize_codebook, dequantizer.py:88-131)
 *   svd_up      rank-r factor, element (n, j) at svd_up[n*up_stride_n + j*up_stride_r]; NULL if none
 *   svd_down    element (j, k) at svd_down[j*down_stride_r + k*down_stride_k]
 *   svd_dtype   SDNQ_BF16 / SDNQ_F16 / SDNQ_F32; the sum is rounded to svd_dtype as addmm_ does
 *   hadamard_group  0 = none, else un-rotate along K in groups of this size (power of two, 4..256)
 *   out         [N,K] row-major, out_dtype SDNQ_BF16 / SDNQ_F16 / SDNQ_F32 */
SDNQ_API int sdnq_b200_dequant(const void* weight, const sdnq_weight_format* fmt, const float* scale, const float* zero_point,
                      int codebook, int64_t N, int64_t K, int64_t group_size,
                      const void* svd_up, int64_t up_stride_n, int64_t up_stride_r,
                      const void* svd_down, int64_t down_stride_r, int64_t down_stride_k,
                      int svd_rank, int svd_dtype, int hadamard_group,
                      void* out, int out_dtype, void* stream);

/* ---- K3 for convolution weights:  dequantize_symmetric / _asymmetric / _codebook as a broadcast over the quantised view
 *      (dequantizer.py:15-131) for layers whose scale is not one value per contiguous K-group: Conv weights reduce over
 *      the input-channel axis only (scale [N,1,kh,kw], grouped [N,C/g,1,kh,kw]; quantizer.py:95-110, 185-199), ConvTranspose
 *      weights over axis 0.
 *   weight         stored codes in flattened order of the quantised view (packed or not)
 *   dims[ndim]     shape of the quantised view (ndim <= 6, product % 8 == 0)
 *   scale_strides  element stride of scale / zero_point (and, times 2^bits, of the codebook levels) along each axis of the
 *                  view, 0 on broadcast axes
 *   addend         optional tensor of the same number of elements added in f32 before the cast: mm(svd_up, svd_down) of an
 *                  SVD layer (dequantizer.py:36-37, 72-73); NULL if none
 *   out            contiguous, dims-shaped, out_dtype SDNQ_BF16 / SDNQ_F16 / SDNQ_F32 */
SDNQ_API int sdnq_b200_dequant_nd(const void* weight, const sdnq_weight_format* fmt, const float* scale, const float* zero_point,
                         int codebook, int ndim, const int64_t* dims, const int64_t* scale_strides,
                         const void* addend, int addend_dtype, void* out, int out_dtype, void* stream);

/* ---- K4 re-quantise for matmul:  SDNQDequantizer.re_quantize_matmul (dequantizer.py:204-239, 353-386)
 * dequant to f32 (no SVD, no un-rotate) then row-wise re-quantise to the matmul dtype.
 *   mm_dtype  SDNQ_I8 (symmetric, quantize_int_mm), SDNQ_U8 (asymmetric int8 codes + zero point,
 *             quantize_uint_mm) or SDNQ_F8E4M3 (quantize_fp_mm)
 *   wq        [N,K] 1-byte codes (the K-major [K,N] operand), sw [N] f32, zw [N] f32 (SDNQ_U8 only)
 *   colsum    optional int32 [N] column sums of wq (needed by the uint8 forward, linear_uint8.py:66-73) */
SDNQ_API int sdnq_b200_requant(const void* weight, const sdnq_weight_format* fmt, const float* scale, const float* zero_point,
                      int codebook, int64_t N, int64_t K, int64_t group_size, int mm_dtype,
                      void* wq, float* sw, float* zw, int32_t* colsum, void* stream);

/* ---- K2 activation pre-pass:  rotate_hadamard + quantize_{int,uint,fp}_mm
 *      (quant_utils.py:193-209, 264-299; linear_int8.py:14-22, 55-56, 63-69)
 * x [M,K] row-major (row stride ldx elements) of x_dtype (SDNQ_BF16 / SDNQ_F16 / SDNQ_F32)
 *   -> optional rotation in groups of hadamard_group (0 = none), rounded back to x_dtype
 *   -> per-row scale sx[M] = amax/127 (I8), (max-min)/255 + zx[M] (U8), amax/448 (F8E4M3)
 *   -> xq [M,K] 1-byte codes
 *   rowsum  optional int32 [M] sum of codes per row (zero-point term); x_rot optional [M,K] x_dtype copy of the
 *           rotated activations (operand of the SVD branch).  K % 8 == 0. */
SDNQ_API int sdnq_b200_act_quant(const void* x, int x_dtype, int64_t M, int64_t K, int64_t ldx, int hadamard_group,
                        int mm_dtype, void* xq, float* sx, float* zx, int32_t* rowsum, void* x_rot, void* stream);

/* ---- K2 for convolutions:  process_conv_input (layers/conv/forward.py:30-76: F.unfold(...).transpose(1, 2)) +
 *      rotate_hadamard + quantize_{int,uint,fp}_mm_input of conv_{int8,uint8,fp8}_matmul (layers/conv/conv_int8.py:17-90).
 * The im2col matrix [M, K] (M = batch * H_out * W_out rows in (b, oh, ow) order, K = channels * kernel_h * kernel_w columns in
 * (c, i, j) order, zero padding) is never written: each row is gathered from the strided input, rotated, row-quantised and
 * only the 1-byte codes xq [M, K] (+ sx / zx / rowsum / optional x_rot, as sdnq_b200_act_quant) are stored.  conv1d: height = 1.
 * K % 8 == 0.  Strides are in elements and must be non-negative. */
typedef struct sdnq_conv2d_geometry {
    int64_t batch, channels, height, width;
    int64_t x_stride_b, x_stride_c, x_stride_h, x_stride_w;
    int32_t kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w;
} sdnq_conv2d_geometry;
SDNQ_API int sdnq_b200_conv_act_quant(const void* x, int x_dtype, const sdnq_conv2d_geometry* geometry, int hadamard_group,
                             int mm_dtype, void* xq, float* sx, float* zx, int32_t* rowsum, void* x_rot, void* stream);

/* The same, with a caller-owned workspace: 1x1 and 3x3 kernels without rotation / x_rot then take the tiled path (a coalesced
 * per-pixel statistics pass into the workspace + a shared-memory tiled quantiser: bit-identical results, ~10x the throughput of
 * the gather kernel); every other case falls through to sdnq_b200_conv_act_quant.  workspace: at least
 * sdnq_b200_conv_act_quant_workspace_bytes() bytes, 8-byte aligned; contents are scratch. */
SDNQ_API size_t sdnq_b200_conv_act_quant_workspace_bytes(const sdnq_conv2d_geometry* geometry, int mm_dtype);
SDNQ_API int sdnq_b200_conv_act_quant_ws(const void* x, int x_dtype, const sdnq_conv2d_geometry* geometry, int hadamard_group,
                                int mm_dtype, void* xq, float* sx, float* zx, int32_t* rowsum, void* x_rot,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ---- conv output layout:  `.view(B, H_out, W_out, N).permute(0, 3, 1, 2).contiguous()` of conv_int8_matmul
 *      (layers/conv/conv_int8.py:83-89).  in [batch*hw, channels] row-major (the GEMM output) -> out [batch, channels, hw];
 *      elem_bytes 2 (bf16 / f16) or 4 (f32). */
SDNQ_API int sdnq_b200_rows_to_nchw(const void* in, void* out, int elem_bytes, int64_t batch, int64_t hw, int64_t channels, void* stream);

/* ---- K1 scaled matmul:  int_scaled_mm_func / fp8_scaled_mm_func (kernel_wrappers.py:193-204) ->
 *      sdnq_scaled_mm (kernels/triton_scaled_mm.py:239-275)
 * out[M,N] = cast( fma( f32(A @ B) * sx[m], sw[n], bias ) )        (no bias: (acc*sx)*sw)
 *   a   [M,K] row-major 1-byte codes (SDNQ_I8 or SDNQ_F8E4M3), row stride K
 *   b   the [K,N] operand stored K-major, i.e. physically [N,K] row-major, same dtype as a
 *       (ab_dtype = SDNQ_F8E5M2: a is float8_e4m3fn, b is float8_e5m2)
 *   sx  f32 [M], sw f32 [N]
 *   bias        NULL, or [N] (bias_ld = 0), or [M,N] (bias_ld = row stride in elements); bias_dtype f32/bf16/f16
 *   zero-point rank-1 terms, all optional (NULL = absent), added to bias in f32 before the fma in the
 *   reference's order (linear_int8.py:65-69, linear_uint8.py:66-73):
 *       rowsum[M] (int32) with zp[N]  :  (f32(rowsum[m]) * sx[m]) * zp[n]
 *       colsum[N] (int32) with zx[M]  :  (f32(colsum[n]) * sw[n]) * zx[m]   and   K * (zx[m] * zp[n])
 *   out  [M,N] row-major, out_dtype SDNQ_BF16 / SDNQ_F16 / SDNQ_F32.  K % 16 == 0, N % 8 == 0.
 *   workspace: none. */
SDNQ_API int sdnq_b200_scaled_mm(const void* a, const void* b, int ab_dtype, const float* sx, const float* sw,
                        const void* bias, int bias_dtype, int64_t bias_ld,
                        const int32_t* rowsum, const float* zp, const int32_t* colsum, const float* zx,
                        void* out, int out_dtype, int64_t M, int64_t N, int64_t K, void* stream);

/* The same scaled matmul with a caller-owned workspace, which lets the library schedule "stream-K": int8 GEMMs whose 128 x 128 tiles
 * would leave a large part of the last wave of SMs idle (SD-XL: 80 tiles on 148 SMs) split their k-blocks evenly over all SMs; partial
 * 32-bit accumulators are parked in the workspace and added by the CTA that finishes the tile -- integer sums, so the output is
 * bit-identical to sdnq_b200_scaled_mm.  The workspace (sdnq_b200_scaled_mm_workspace_bytes() bytes, 16-byte aligned) must be
 * ZERO before its first use and belong to one stream at a time; the library leaves its flag area zero after every launch.
 * workspace = NULL (or too small): exactly sdnq_b200_scaled_mm. */
SDNQ_API size_t sdnq_b200_scaled_mm_workspace_bytes(void);
SDNQ_API int sdnq_b200_scaled_mm_ws(const void* a, const void* b, int ab_dtype, const float* sx, const float* sw,
                        const void* bias, int bias_dtype, int64_t bias_ld,
                        const int32_t* rowsum, const float* zp, const int32_t* colsum, const float* zx,
                        void* out, int out_dtype, int64_t M, int64_t N, int64_t K,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- K1 with the packed-weight unpack fused into the GEMM (north_star "unpack + scale in the GEMM prologue"):
 *      linear_int8.py:38-44 / linear_fp8.py:38 (per-call unpack_int / unpack_float(...).t_() of a row-wise packed weight,
 *      packed_int/unpack.py:233-372, packed_float.py:85-132) + the scaled matmul above.
 * b_packed  [N, K*bits/8] row-wise packed storage exactly as the reference keeps it (packed_int/pack.py: octets of 8 codes in
 *           `bits` bytes), bits = 2..7:
 *             integer codes (int2..int7 / uint2..uint7)  with int8 activations  -> expanded to int8 (code - 2^(bits-1) for signed formats),
 *             minifloat codes eXmY with X <= 4, Y <= 3   with e4m3 activations  -> expanded to the e4m3 byte of the same value.
 *           TMA stages the packed tile in shared memory, four unpack warps expand it in the UMMA layout, the row-wise scale sw[n]
 *           (and, for unsigned integers, the zero-point term rowsum[m]*sx[m]*zp[n]) is applied in the epilogue.  The result is, bit
 *           for bit, sdnq_b200_unpack followed by sdnq_b200_scaled_mm -- without the N*K-byte expanded copy.
 * a, sx, bias, out as in sdnq_b200_scaled_mm.  K * bits % 128 == 0 (16-byte row pitch). */
SDNQ_API int sdnq_b200_scaled_mm_packed(const void* a, const void* b_packed, const sdnq_weight_format* b_fmt, const float* sx,
                                        const float* sw, const void* bias, int bias_dtype, int64_t bias_ld,
                                        const int32_t* rowsum, const float* zp, void* out, int out_dtype,
                                        int64_t M, int64_t N, int64_t K, void* stream);

/* ---- plain matmul:  int_mm_func / fp8_mm_func (kernel_wrappers.py:160-181) -> sdnq_triton_mm
 *      (kernels/triton_mm.py:119-150).  out[M,N] = A @ B as int32 (I8) or f32 (F8E4M3). */
SDNQ_API int sdnq_b200_mm(const void* a, const void* b, int ab_dtype, void* out, int64_t M, int64_t N, int64_t K, void* stream);

/* ---- fused W8A8 Linear:  quantized_linear_forward_{int8,uint8,fp8}_matmul for layers whose stored
 *      weight is already the matmul operand (linear_int8.py:100-125 with re_quantize_for_matmul cached).
 * One call = K2 + K1 on `stream` (programmatic dependent launch between them); with SDNQ_B200_FUSED=1 in the
 * environment the single-launch kernel below is used wherever it applies.  The workspace holds the strip counters of
 * the fused kernel, xq, sx, zx and rowsum.  Its first 4096 bytes must be zero when the workspace is used for the first
 * time (zero-fill it once after allocating); every call leaves them zero again.  One workspace per stream: calls that
 * may run concurrently need separate workspaces.
 *   mm_dtype SDNQ_I8 / SDNQ_U8 / SDNQ_F8E4M3 / SDNQ_F8E5M2 (e4m3 activation codes x stored float8_e5m2 weight);
 *   wq physical [N,K]; zp / colsum as in scaled_mm. */
SDNQ_API size_t sdnq_b200_linear_w8a8_workspace_bytes(int64_t M, int64_t K);
SDNQ_API int sdnq_b200_linear_w8a8(const void* x, int x_dtype, int64_t ldx, const void* wq, int mm_dtype,
                          const float* sw, const float* zp, const int32_t* colsum,
                          const void* bias, int bias_dtype, int hadamard_group,
                          void* out, int out_dtype, int64_t M, int64_t N, int64_t K,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- K3, batched: the weights of SEVERAL layers dequantised by one persistent launch.  The reference dequantises inside every
 *      forward (layers/linear/forward.py:24-26 -> SDNQDequantizer.__call__, dequantizer.py:389-429); the weights are frozen, so the
 *      host layer learns the order in which layers are called and dequantises the next few layers' weights together, ahead of
 *      their GEMMs, on a side stream.  Every weight is still dequantised exactly once per use, with the arithmetic of
 *      sdnq_b200_dequant (same kernel, a table of weights instead of one).
 *   jobs         per weight: the arguments of sdnq_b200_dequant (no codebook, no Hadamard un-rotate) plus its output
 *   host_table   sdnq_b200_dequant_batch_table_bytes(n_jobs) bytes of HOST memory the plan is written to; copy it to device memory
 *                (128-byte aligned) once and pass that copy to every run.  The plan embeds the jobs' pointers: it stays valid
 *                while those allocations do.
 *   info         int32[4] filled by the plan, passed back to run
 * Covers 4-bit integer weights with bf16 SVD factors (rank 16 / 32 / 64, stored K-major) and bf16 output; anything else: SDNQ_EUNSUPPORTED. */
typedef struct sdnq_dequant_job {
    const void* weight;
    sdnq_weight_format fmt;
    const float* scale;
    const float* zero_point;
    int64_t N, K, group_size;
    const void* svd_up;
    int64_t up_stride_n, up_stride_r;
    const void* svd_down;
    int64_t down_stride_r, down_stride_k;
    int32_t svd_rank, svd_dtype;
    void* out;
    int32_t out_dtype, reserved;
} sdnq_dequant_job;
SDNQ_API size_t sdnq_b200_dequant_batch_table_bytes(int n_jobs);
SDNQ_API int sdnq_b200_dequant_batch_plan(const sdnq_dequant_job* jobs, int n_jobs, void* host_table, int32_t* info);
SDNQ_API int sdnq_b200_dequant_batch_run(const void* device_table, const int32_t* info, void* stream);

/* ---- quantized embedding lookup (quantized_embedding / quantized_embedding_forward, layers/embedding/forward.py:14-104): the reference
 *      unpacks the WHOLE table, indexes weight / scale / zero_point / svd_up with the token ids and dequantises the selected rows.
 *      Here the dequant kernel reads stored row indices[n] for output row n directly (the table is never unpacked).
 *   weight, fmt, scale, zero_point, codebook, group_size, svd_*, hadamard_group   the stored [V, D] table, as for sdnq_b200_dequant
 *   indices      int64 [n_indices] on the device; an id outside [0, V) traps (torch raises a device-side assert)
 *   embed_scale  result.mul_(embed_scale) on the rounded rows (1.0f: none)
 *   out          [n_indices, D] row-major of out_dtype */
SDNQ_API int sdnq_b200_embedding(const void* weight, const sdnq_weight_format* fmt, const float* scale, const float* zero_point, int codebook,
                                 int64_t V, int64_t D, int64_t group_size, const void* svd_up, int64_t up_stride_n, int64_t up_stride_r,
                                 const void* svd_down, int64_t down_stride_r, int64_t down_stride_k, int svd_rank, int svd_dtype,
                                 int hadamard_group, const int64_t* indices, int64_t n_indices, float embed_scale, void* out, int out_dtype,
                                 void* stream);

/* ---- K8 load-time quantisation of a weight: scale, round, clamp and pack in one pass (the middle of sdnq_quantize_layer_weight,
 *      quantizer.py:236-253: quantize_weight, quant_utils.py:27-56, then pack_int, packed_int/__init__.py:76-80 + pack.py:201-321,
 *      or pack_float, packed_float.py:26-82).
 *   w            [N,K] row-major f32 / bf16 / f16 (after the optional Hadamard rotation / SVD subtraction, which stay host-side);
 *                K % 8 == 0, 16-byte aligned
 *   group_size   scale groups of that many consecutive weights along K (a multiple of 8 dividing K); <= 0 or >= K: row-wise
 *   fmt          integer formats of 2..8 bits: symmetric  scale = amax / qmax,  q = round_half_even(w / scale)  for signed formats,
 *                asymmetric  scale = (max - min) / qmax, zero_point = min,  q = round_half_even((w - zero_point) / scale)  for unsigned
 *                ones; true f32 divisions, clamp to the format's range -- the arithmetic of the reference on the CPU, bit for bit.
 *                Float formats of 2..8 bits (SDNQ_W_FP8_E4M3FN, SDNQ_W_FP8_E5M2, SDNQ_W_MINIFLOAT signed "fn" / unsigned "fnu"):
 *                the same scales over the format's largest value, q = nan_to_num(w / scale) clamped to the range, then the cast
 *                to torch.float8_* (round to nearest even) or pack_float's own bit arithmetic (packed_float.py:26-82: normals round
 *                up when the top four dropped mantissa bits exceed one half, subnormals are round_half_even(|q| * 2^M / min_normal))
 *   scale_dtype  SDNQ_F32, or SDNQ_BF16 / SDNQ_F16: scale and zero point are rounded to that type before the division (dequantize_fp32=False)
 *   codes        bits < 8: N*K*bits/8 packed bytes (signed integer codes offset-binary);  bits == 8: N*K one-byte codes (int8 two's
 *                complement / uint8 / float8 bit patterns / 8-bit minifloat codes)
 *   scale, zero_point   f32 [N * K / group_size]  (zero_point written for unsigned formats only) */
SDNQ_API int sdnq_b200_quantize_weight(const void* w, int w_dtype, int64_t N, int64_t K, int64_t group_size, const sdnq_weight_format* fmt,
                                       int scale_dtype, void* codes, float* scale, float* zero_point, void* stream);

/* ---- K5 small-M Linear on 8-bit weights ("W8A16 GEMV"):  the rows < 32 branch of every quantized-matmul forward
 *      (linear_int8.py:102-103, linear_uint8.py:107-108, linear_fp8.py:83-84: dequantise the weight, then F.linear) without
 *      materialising the dequantised weight:
 *          out[m,n] = sw[n] * sum_k x[m,k] * q[n,k]  (+ zp[n] * sum_k x[m,k])  + bias[n]
 *   x     [M,K] bf16 / f16, row stride ldx (for use_hadamard layers: already rotated, e.g. the x_rot output of act_quant)
 *   wq    physical [N,K] 1-byte codes, w_dtype SDNQ_I8, SDNQ_F8E4M3 or SDNQ_F8E5M2 (row-wise scales sw[N], optional zero points zp[N])
 *   out   [M,N] row-major in x_dtype.   1 <= M <= 32, K % 16 == 0.
 * The codes are read once (N*K bytes instead of the 3*N*K of dequantise-then-GEMM); accumulation is f32 on the tensor cores. */
SDNQ_API int sdnq_b200_linear_small_m(const void* x, int x_dtype, int64_t ldx, const void* wq, int w_dtype, const float* sw,
                             const float* zp, const void* bias, int bias_dtype, void* out,
                             int64_t M, int64_t N, int64_t K, void* stream);

/* ---- K5p small-M Linear straight from the STORED weight (packed sub-byte integers, minifloats, fp8 / int8; row-wise or
 *      group-wise scales; optional zero points): the rows < 32 branch (layers/linear/forward.py:24-26 with
 *      use_quantized_matmul=False, linear_int8.py:102-103 otherwise: SDNQDequantizer.__call__ then F.linear) without writing the
 *      dequantised weight:   out[m,n] = sum_k x[m,k] * cast_T(q[n,k] * s[n,k/g] (+ zp[n,k/g]))  + bias
 *   weight      the stored tensor as the reference keeps it: packed along the flattened [N,K] weight (packed_int/pack.py), or
 *               plain 1-byte codes [N,K];  fmt as for sdnq_b200_unpack (1..8 bits, word_bytes 1)
 *   scale / zero_point   f32 [N, K/group_size] (zero_point NULL for symmetric formats); group_size <= 0 or >= K: row-wise;
 *               otherwise a multiple of 8 dividing K.  Every weight is rounded to the activation dtype before the product,
 *               exactly the value the reference's dequantised weight holds (dequantizer.py:15-84).
 *   bias        NULL, [N] (bias_ld = 0) or [M,N] (bias_ld = row stride; e.g. the SVD term (x @ svd_down^T) @ svd_up^T + bias)
 *   x, out      [M,K] / [M,N] bf16 or f16 (x_dtype), 1 <= M <= 32, K % 16 == 0, ldx % 8 == 0.  Accumulation f32 on the tensor cores. */
SDNQ_API int sdnq_b200_linear_small_m_packed(const void* x, int x_dtype, int64_t ldx, const void* weight, const sdnq_weight_format* fmt,
                                             const float* scale, const float* zero_point, int64_t group_size, const void* bias,
                                             int bias_dtype, int64_t bias_ld, void* out, int64_t M, int64_t N, int64_t K, void* stream);

/* ---- K6 dequant-path Linear in one launch ("W4A16"):  quantized_linear_forward (layers/linear/forward.py:24-26) =
 *      SDNQDequantizer.__call__ (dequantizer.py:389-429) + F.linear, for use_quantized_matmul=False layers with 4-bit weights:
 *          out[m,n] = sum_k x[m,k] * cast_T(q[n,k] * s[n,k/g] (+ zp[n,k/g]))  +  sum_j cast_T(sum_k x[m,k] down[j,k]) * up[n,j]  + bias[n]
 *      The packed codes are staged by TMA and dequantised (unpack, group scale, zero point, rounding to the activation dtype T --
 *      exactly the values dequantize_symmetric / _asymmetric produce before the SVD add, dequantizer.py:15-84) into the UMMA
 *      operand ring by the GEMM's prologue warps; the contraction runs on tcgen05 kind::f16 with f32 accumulation; the SVD
 *      correction W += svd_up @ svd_down (dequantizer.py:69-79) is applied as a rank-r second accumulate on the activations,
 *      (x down^T) up^T, so the [N,K] weight is never written in 16 bits.
 *   x          [M,K] bf16 / f16 (x_dtype), row stride ldx (use_hadamard layers: pass the rotated activations, x_rot of act_quant)
 *   weight     int4 / uint4 codes of the [N,K] weight packed as packed_int/pack.py:273-276 lays them out ([N*K/2] bytes)
 *   scale / zero_point   f32 [N, K/group_size]; group_size <= 0 or >= K: row-wise, else a multiple of 32 dividing K
 *   svd_down_rk [svd_rank, K] row-major and svd_up_nr [N, svd_rank] row-major, both of x_dtype (svd_rank 0: no SVD term; else 16 / 32 / 64)
 *   bias       NULL or [N];   out [M,N] row-major of x_dtype.   K % 64 == 0, N % 8 == 0, ldx % 8 == 0. */
SDNQ_API int sdnq_b200_linear_w4a16(const void* x, int x_dtype, int64_t ldx, const void* weight, const sdnq_weight_format* fmt,
                                    const float* scale, const float* zero_point, int64_t group_size,
                                    const void* svd_down_rk, const void* svd_up_nr, int svd_rank,
                                    const void* bias, int bias_dtype, void* out, int64_t M, int64_t N, int64_t K, void* stream);

/* ---- K7 + K1: the SVD branch of the W8A8 forwards (get_int8_matmul_inputs / get_uint8_matmul_inputs / get_fp8_matmul_inputs with
 *      svd_up / svd_down, layers/linear/linear_int8.py:57-62, linear_uint8.py:58-63, linear_fp8.py:47-52, conv_int8.py:56-61):
 *          bias2d = bias + torch.mm(torch.mm(x_rot, svd_down), svd_up)          (reference: two library GEMMs + a dense [M,N] bias)
 *      Here:  low = sdnq_b200_svd_low(x_rot, svd_down)  [M, r], rounded to the activation dtype exactly where torch.mm rounds it, then
 *      sdnq_b200_scaled_mm_svd accumulates low @ svd_up as a rank-r tcgen05.mma (kind::f16, f32 accumulator in its own TMEM region)
 *      per output tile and its epilogue adds it to the bias in f32 -- no [M,N] bias tensor, no library GEMM.
 *   x            [M,K] bf16 / f16 (the rotated, un-quantised activations: x_rot of act_quant, or x itself without Hadamard), row stride ldx
 *   svd_down_rk  [svd_rank, K] row-major of x_dtype (the stored [K, r] factor transposed once);  svd_rank % 8 == 0, <= 64;  K % 16 == 0
 *   low          [M, svd_rank] row-major of x_dtype (written) */
SDNQ_API int sdnq_b200_svd_low(const void* x, int x_dtype, int64_t ldx, const void* svd_down_rk, int svd_rank, void* low,
                               int64_t M, int64_t K, void* stream);

/*   a, b, sx, sw, bias, rowsum, zp, colsum, zx, out, M, N, K   as for sdnq_b200_scaled_mm (bias_ld != 0: an [M,N] bias on top)
 *   b_fmt        NULL: b is the physical [N,K] 1-byte operand;  else int4 / uint4: b is the stored packed weight (as sdnq_b200_scaled_mm_packed)
 *   svd_low      [M, svd_rank] from sdnq_b200_svd_low;  svd_up_nr [N, svd_rank] row-major;  both svd_dtype (SDNQ_BF16 / SDNQ_F16)
 *   svd_rank     16, 32 or 64 (pad the factors with zero columns for other ranks);  out_dtype SDNQ_BF16 / SDNQ_F16 */
SDNQ_API int sdnq_b200_scaled_mm_svd(const void* a, const void* b, int ab_dtype, const sdnq_weight_format* b_fmt, const float* sx, const float* sw,
                                     const void* bias, int bias_dtype, int64_t bias_ld, const int32_t* rowsum, const float* zp,
                                     const int32_t* colsum, const float* zx, const void* svd_low, const void* svd_up_nr, int svd_rank,
                                     int svd_dtype, void* out, int out_dtype, int64_t M, int64_t N, int64_t K, void* stream);

/* ---- K1, grouped: several Linears that read the SAME activations (to_q / to_k / to_v of an attention block, the cross-attention
 *      to_k / to_v pair, gate / up projections) as ONE launch.  The reference runs one scaled matmul per Linear
 *      (linear_int8.py:100-125 each time); here the siblings' matmul operands are concatenated along N once, the quantised
 *      activations are read by one persistent grid, and every sibling gets its own contiguous output -- per element exactly the
 *      value sdnq_b200_scaled_mm / _packed produces for that sibling alone.
 *   b_cat        the siblings' physical [N_g, K] operands (or, b_fmt = int4 / uint4: their packed rows) stacked; segment g starts at row
 *                seg_start[g] (a multiple of 128; all multiples of 256 lets the kernel use 256-wide tiles and CTA pairs) and holds
 *                seg_n[g] rows followed by zero rows up to seg_start[g+1];  seg_start has n_groups + 1 entries, the last = total rows
 *   sw_cat, bias_cat, zp_cat, colsum_cat   the per-output-channel vectors laid out the same way (bias_cat / zp_cat / colsum_cat may be NULL)
 *   outs[g]      [M, seg_n[g]] row-major of out_dtype;  1 <= n_groups <= 8;  other arguments as for sdnq_b200_scaled_mm */
SDNQ_API int sdnq_b200_scaled_mm_grouped(const void* a, const void* b_cat, int ab_dtype, const sdnq_weight_format* b_fmt, const float* sx,
                                         const float* sw_cat, const void* bias_cat, int bias_dtype, const int32_t* rowsum, const float* zp_cat,
                                         const int32_t* colsum_cat, const float* zx, int n_groups, const int64_t* seg_start, const int64_t* seg_n,
                                         void* const* outs, int out_dtype, int64_t M, int64_t K, void* stream);

/* The same Linear as ONE kernel launch: the GEMM kernel row-quantises the activations itself (every CTA takes a share
 * of the rows: bulk copy to shared memory, warp-reduction amax, quantise, codes + scales to the workspace) and its TMA
 * producers pick the quantised strips up through release/acquire strip counters -- linear_int8.py:14-22 + 100-125 /
 * linear_fp8.py:14-22 + 81-104 in one kernel.  Covers symmetric int8 / float8_e4m3fn without Hadamard or zero-points,
 * x and out both bf16 or both f16, vector (or no) bias; anything else returns SDNQ_EUNSUPPORTED.  Bit-identical to
 * act_quant + scaled_mm.  Workspace contract as for sdnq_b200_linear_w8a8. */
SDNQ_API int sdnq_b200_linear_w8a8_fused(const void* x, int x_dtype, int64_t ldx, const void* wq, int mm_dtype,
                                const float* sw, const void* bias, int bias_dtype,
                                void* out, int out_dtype, int64_t M, int64_t N, int64_t K,
                                void* workspace, size_t workspace_bytes, void* stream);

/* 0 if `stream` is not being captured into a CUDA graph, else the (non-zero) id of the capture; -1 on error.  The host layer
 * keys everything it caches across calls (quantised activations shared by sibling projections) on it, so that nothing computed
 * eagerly is baked into a graph and nothing captured in one graph is taken for valid in another. */
SDNQ_API int64_t sdnq_b200_stream_capture_id(void* stream);

/* number of kernels launched by this library on the calling thread since the last reset
 * (bench.py reports it as gpu_launches) */
SDNQ_API int64_t sdnq_b200_launch_count(int reset);

/* ---- K9 quantized attention forward:  sdnq_attn_kernel (kernels/triton_atten.py:143-335) as launched by sdnq_atten_fwd
 *      (:338-386) on the operands quantize_attn (:443-487) produces.
 *   q [Z,H,QN,HD], k [Z,KH,KN,HD]   contiguous 1-byte codes, qk_dtype SDNQ_I8 or SDNQ_F8E4M3; HD % 16 == 0, HD <= 128
 *   q_scale [Z,H,QN], k_scale [Z,KH,KN]   f32 per-row scales (quantize_int_mm / quantize_fp_mm over the head dim)
 *   v [Z,VH,KN,HDV]                 contiguous SDNQ_BF16 / SDNQ_F16 values with v_scale = NULL (the unquantised P.V of
 *                                   pv_matmul_dtype = None, :319-321), or SDNQ_I8 / SDNQ_F8E4M3 codes with per-key scales
 *                                   v_scale [Z,VH,KN] (quantised P.V, :298-318: P is scaled by v_scale, quantised per row and key
 *                                   tile, and its row scale applied to the product); HDV 64 or 128
 *   mask                            NULL, or SDNQ_I8 (boolean: 0 = masked out) / SDNQ_F32 (additive) with 4 element strides over
 *                                   (z, h, q, k), 0 on broadcast axes (:370-376)
 *   out [Z,H,QN,HDV], lse [Z,H,QN] (optional: m + log2(l), :326-332) of out_dtype SDNQ_BF16 / SDNQ_F16 / SDNQ_F32
 *   heads of k / v map as h * KH / H and h * VH / H (:197-198); is_causal aligns the diagonal top-left (:277-278)
 *   workspace   sdnq_b200_attention_workspace_bytes(Z, VH, KN, HDV) bytes, 16-byte aligned: V^T written by a pre-pass kernel */
SDNQ_API size_t sdnq_b200_attention_workspace_bytes(int64_t Z, int64_t VH, int64_t KN, int64_t HDV);
SDNQ_API int sdnq_b200_attention(const void* q, const void* k, const void* v, int qk_dtype, int v_dtype, const float* q_scale,
                        const float* k_scale, const float* v_scale, const void* mask, int mask_dtype, const int64_t* mask_strides, void* out,
                        void* lse, int out_dtype, int64_t Z, int64_t H, int64_t KH, int64_t VH, int64_t QN, int64_t KN,
                        int64_t HD, int64_t HDV, float sm_scale, int is_causal, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ---- smooth-K of quantize_attn (kernels/triton_atten.py:456-461): out = k.to(f32) - mean(k, tokens) per (batch, head, channel)
 *   k [heads, N, HD] contiguous of k_dtype (SDNQ_BF16 / SDNQ_F16 / SDNQ_F32) -> out of out_dtype (SDNQ_F32, or k_dtype's 16-bit
 *   type when the Hadamard rotation follows in that type, :463-466); HD <= 256 */
SDNQ_API int sdnq_b200_smooth_k(const void* k, int k_dtype, int64_t heads, int64_t N, int64_t HD, void* out, int out_dtype, void* stream);

/* ---- attention operand pre-pass without a rotation (quantize_attn, kernels/triton_atten.py:456-471):
 * sdnq_b200_attn_colmean:  k [heads, N, HD] -> mean [heads, HD] f32, the per-channel token means smooth-K subtracts (:456-461)
 * sdnq_b200_attn_quant:    x [rows, HD] (SDNQ_BF16 / SDNQ_F16 / SDNQ_F32), optionally minus mean[row / rows_per_head] in f32,
 *                          -> per-row scale[rows] = amax / 127 (SDNQ_I8) or / 448 (SDNQ_F8E4M3) and 1-byte codes xq [rows, HD]
 *                          (quantize_int_mm / quantize_fp_mm over the head dim, quant_utils.py:264-299; same exact division as K2)
 * HD a power of two, 16 .. 256 (get_attn_inputs pads to one, :514-519) */
SDNQ_API int sdnq_b200_attn_colmean(const void* k, int k_dtype, int64_t heads, int64_t N, int64_t HD, float* mean, void* stream);
SDNQ_API int sdnq_b200_attn_quant(const void* x, int x_dtype, int64_t rows, int64_t HD, const float* mean, int64_t rows_per_head,
                         int mm_dtype, void* xq, float* scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDNQ_B200_H */
