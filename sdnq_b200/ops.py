"""torch.Tensor-facing wrappers over the C ABI.  PyTorch is only the allocator / stream provider here: every
function resolves raw device pointers and calls the hand-written sm_100a kernels; nothing computes in eager."""
import torch

from . import _lib
from ._lib import (SDNQ_BF16, SDNQ_F16, SDNQ_F32, SDNQ_F8E4M3, SDNQ_F8E5M2, SDNQ_I8, SDNQ_I32, SDNQ_U8, SDNQ_W_FP8_E4M3FN, SDNQ_W_FP8_E5M2,
                   SDNQ_W_INT, SDNQ_W_MINIFLOAT, WeightFormat, check)
from .common import dtype_dict

_TORCH_TO_CODE = {torch.float32: SDNQ_F32, torch.bfloat16: SDNQ_BF16, torch.float16: SDNQ_F16, torch.int8: SDNQ_I8,
                  torch.uint8: SDNQ_U8, torch.float8_e4m3fn: SDNQ_F8E4M3, torch.int32: SDNQ_I32, torch.float8_e5m2: SDNQ_F8E5M2}
_MM_CODE = {"int8": SDNQ_I8, "uint8": SDNQ_U8, "float8_e4m3fn": SDNQ_F8E4M3, "fp8": SDNQ_F8E4M3}
_MM_TORCH = {SDNQ_I8: torch.int8, SDNQ_U8: torch.int8, SDNQ_F8E4M3: torch.float8_e4m3fn}


def dtype_code(dtype: torch.dtype) -> int:
    try:
        return _TORCH_TO_CODE[dtype]
    except KeyError:
        raise _lib.SDNQKernelError(f"dtype {dtype} is not supported by the sdnq_b200 kernels") from None


def mm_code(matmul_dtype: str) -> int:
    try:
        return _MM_CODE[matmul_dtype]
    except KeyError:
        raise _lib.SDNQKernelError(f"quantized_matmul_dtype {matmul_dtype!r} has no sm_100a kernel (int8, uint8, float8_e4m3fn do)") from None


def weight_format(weights_dtype: str, storage: torch.Tensor | None = None) -> WeightFormat:
    """One row of `dtype_dict` -> the C ABI's sdnq_weight_format."""
    e = dtype_dict[weights_dtype]
    if e["num_bits"] > 8:
        raise _lib.SDNQKernelError(f"weights_dtype {weights_dtype!r}: formats wider than 8 bits have no CUDA kernel yet")
    word_bytes = 1
    if e["num_bits"] == 1 and storage is not None:
        word_bytes = storage.element_size()      # upstream stores uint1 as one int64 per packed byte
    if e["is_integer"]:
        return WeightFormat(SDNQ_W_INT, e["num_bits"], int(e["is_unsigned"]), 0, 0, word_bytes)
    if e["torch_dtype"] == torch.float8_e4m3fn:
        return WeightFormat(SDNQ_W_FP8_E4M3FN, 8, 0, 4, 3, 1)
    if e["torch_dtype"] == torch.float8_e5m2:
        return WeightFormat(SDNQ_W_FP8_E5M2, 8, 0, 5, 2, 1)
    if not e["is_packed"]:
        raise _lib.SDNQKernelError(f"weights_dtype {weights_dtype!r} is not a quantised storage format")
    return WeightFormat(SDNQ_W_MINIFLOAT, e["num_bits"], int(e["is_unsigned"]), e["exponent"], e["mantissa"], word_bytes)


def _operand_code(a_dtype: torch.dtype, b: torch.Tensor) -> int:
    """The C ABI's operand code of a GEMM over activation codes of `a_dtype` and the 1-byte weight `b`.  The weight's own dtype
    decides how its bytes are decoded (a float8_e5m2 weight under a float8_e4m3fn matmul is the mixed pair of torch._scaled_mm);
    anything the tensor cores cannot take raises instead of being reinterpreted."""
    if a_dtype == torch.float8_e4m3fn:
        if b.dtype == torch.float8_e4m3fn:
            return SDNQ_F8E4M3
        if b.dtype == torch.float8_e5m2:
            return SDNQ_F8E5M2
    elif a_dtype == torch.int8 and b.dtype == torch.int8:
        return SDNQ_I8
    raise _lib.SDNQKernelError(f"no sm_100a GEMM for {a_dtype} activation codes x {b.dtype} weight codes")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def capture_id(stream: int) -> int:
    """0 when `stream` (a raw cudaStream_t) is not being captured into a CUDA graph, else the id of the capture."""
    return int(_lib.load().sdnq_b200_stream_capture_id(stream))


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.SDNQKernelError("sdnq_b200 kernels need CUDA tensors (there is no CPU path)")


def physical_nk(t: torch.Tensor) -> torch.Tensor:
    """A reference tensor that is logically [K,N] but K-major (stride (1,K)), or logically [N,K] contiguous, viewed as the
    contiguous [N,K] array that is actually in memory.  Copies only if it is neither."""
    if t.ndim != 2:
        return t.contiguous()
    if t.is_contiguous():
        return t
    if t.t().is_contiguous():
        return t.t()
    return t.contiguous()


def unpack(packed: torch.Tensor, weights_dtype: str, shape, dtype: torch.dtype | None = None) -> torch.Tensor:
    """packed_int.unpack_int / packed_float.unpack_float on device."""
    _require_cuda(packed)
    e = dtype_dict[weights_dtype]
    if dtype is None:
        dtype = (torch.uint8 if e["is_unsigned"] else torch.int8) if e["is_integer"] else torch.float32
    packed = packed.contiguous()
    out = torch.empty(tuple(shape), dtype=dtype, device=packed.device)
    fmt = weight_format(weights_dtype, packed)
    with torch.cuda.device(packed.device):
        check(_lib.load().sdnq_b200_unpack(_ptr(packed), fmt, _ptr(out), dtype_code(dtype), out.numel(), _stream(packed)))
    return out


def _group_args(scale: torch.Tensor, N: int, K: int, group_size: int, use_codebook: bool, bits: int):
    if group_size == -2:
        return -2
    per_row = scale.numel() // N // ((1 << bits) if use_codebook else 1)
    return K // max(per_row, 1)


def dequant(weight, weights_dtype, scale, zero_point, N, K, group_size, out_dtype, svd_up=None, svd_down=None,
            svd_layout_matmul=False, hadamard_group=0, use_codebook=False) -> torch.Tensor:
    """K3.  `weight` is the stored tensor (packed 1-D/2-D, or unpacked [N,K] / K-major [K,N]); returns W[N,K] of out_dtype.
    svd factors are passed as stored: matmul layout has svd_up [r,N], svd_down [K,r]; plain layout svd_up [N,r], svd_down [r,K]."""
    _require_cuda(weight, scale)
    lib = _lib.load()
    e = dtype_dict[weights_dtype]
    w = weight if e["is_packed"] else physical_nk(weight)
    w = w.contiguous() if e["is_packed"] else w
    scale = scale.to(torch.float32).contiguous() if scale.dtype != torch.float32 or not scale.is_contiguous() else scale
    if zero_point is not None and (zero_point.dtype != torch.float32 or not zero_point.is_contiguous()):
        zero_point = zero_point.to(torch.float32).contiguous()
    fmt = weight_format(weights_dtype, w)
    gs = _group_args(scale, N, K, group_size, use_codebook, e["num_bits"])
    out = torch.empty((N, K), dtype=out_dtype, device=w.device)
    up_args = (None, 0, 0)
    down_args = (None, 0, 0)
    rank, svd_code = 0, SDNQ_BF16
    if svd_up is not None:
        if svd_layout_matmul:   # svd_up [r,N], svd_down [K,r]
            rank = svd_up.shape[0]
            up_args = (_ptr(svd_up), svd_up.stride(1), svd_up.stride(0))
            down_args = (_ptr(svd_down), svd_down.stride(1), svd_down.stride(0))
        else:                   # svd_up [N,r], svd_down [r,K]
            rank = svd_up.shape[1]
            up_args = (_ptr(svd_up), svd_up.stride(0), svd_up.stride(1))
            down_args = (_ptr(svd_down), svd_down.stride(0), svd_down.stride(1))
        svd_code = dtype_code(svd_up.dtype)
    with torch.cuda.device(w.device):
        check(lib.sdnq_b200_dequant(_ptr(w), fmt, _ptr(scale), _ptr(zero_point), int(use_codebook), N, K, gs,
                                    *up_args, *down_args, rank, svd_code, int(hadamard_group), _ptr(out), dtype_code(out_dtype), _stream(w)))
    return out


def embedding(weight, weights_dtype, scale, zero_point, V, D, group_size, indices: torch.Tensor, out_dtype, svd_up=None, svd_down=None,
              hadamard_group=0, use_codebook=False, embed_scale: float = 1.0) -> torch.Tensor:
    """Quantized embedding lookup: rows `indices` (any shape, integer) of the stored [V, D] table, dequantised -> [*indices.shape, D].
    Arguments as for `dequant` (svd factors in the plain layout: svd_up [V,r], svd_down [r,D])."""
    import ctypes
    _require_cuda(weight, scale, indices)
    e = dtype_dict[weights_dtype]
    w = weight.contiguous() if e["is_packed"] else physical_nk(weight)
    scale = scale.to(torch.float32).contiguous() if scale.dtype != torch.float32 or not scale.is_contiguous() else scale
    if zero_point is not None and (zero_point.dtype != torch.float32 or not zero_point.is_contiguous()):
        zero_point = zero_point.to(torch.float32).contiguous()
    idx = indices.reshape(-1).to(torch.int64).contiguous()
    out = torch.empty((idx.numel(), D), dtype=out_dtype, device=w.device)
    up_args, down_args, rank, svd_code = (None, 0, 0), (None, 0, 0), 0, SDNQ_BF16
    if svd_up is not None:
        rank = svd_up.shape[1]
        up_args = (_ptr(svd_up), svd_up.stride(0), svd_up.stride(1))
        down_args = (_ptr(svd_down), svd_down.stride(0), svd_down.stride(1))
        svd_code = dtype_code(svd_up.dtype)
    gs = _group_args(scale, V, D, group_size, use_codebook, e["num_bits"])
    with torch.cuda.device(w.device):
        check(_lib.load().sdnq_b200_embedding(_ptr(w), weight_format(weights_dtype, w), _ptr(scale), _ptr(zero_point), int(use_codebook), V, D, gs,
                                              *up_args, *down_args, rank, svd_code, int(hadamard_group),
                                              ctypes.cast(idx.data_ptr(), ctypes.POINTER(ctypes.c_int64)), idx.numel(), float(embed_scale),
                                              _ptr(out), dtype_code(out_dtype), _stream(w)))
    return out.view(*indices.shape, D)


def smooth_k(k: torch.Tensor, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """k [..., N, HD] -> k.to(f32) - mean over the N tokens (quantize_attn, kernels/triton_atten.py:456-461), in out_dtype."""
    _require_cuda(k)
    k = k.contiguous()
    N, HD = k.shape[-2], k.shape[-1]
    heads = k.numel() // (N * HD) if k.numel() else 0
    out = torch.empty(k.shape, dtype=out_dtype, device=k.device)
    with torch.cuda.device(k.device):
        check(_lib.load().sdnq_b200_smooth_k(_ptr(k), dtype_code(k.dtype), heads, N, HD, _ptr(out), dtype_code(out_dtype), _stream(k)))
    return out


def attn_quant(x: torch.Tensor, matmul_dtype: str, smooth: bool = False):
    """Attention operand pre-pass without a rotation: x [..., N, HD] -> (codes [..., N, HD], scale [..., N]); `smooth`: subtract the
    per-(batch, head, channel) token means in f32 first (smooth-K).  quantize_attn, kernels/triton_atten.py:456-471."""
    _require_cuda(x)
    x = x.contiguous()
    N, HD = x.shape[-2], x.shape[-1]
    rows = x.numel() // HD
    code = mm_code(matmul_dtype)
    lib = _lib.load()
    xq = torch.empty(x.shape, dtype=_MM_TORCH[code], device=x.device)
    scale = torch.empty(x.shape[:-1], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        mean = None
        if smooth:
            mean = torch.empty((rows // N, HD), dtype=torch.float32, device=x.device)
            check(lib.sdnq_b200_attn_colmean(_ptr(x), dtype_code(x.dtype), rows // N, N, HD, _ptr(mean), _stream(x)))
        check(lib.sdnq_b200_attn_quant(_ptr(x), dtype_code(x.dtype), rows, HD, _ptr(mean), N, code, _ptr(xq), _ptr(scale), _stream(x)))
    return xq, scale


def attention_fwd(q_q: torch.Tensor, k_q: torch.Tensor, v: torch.Tensor, q_scale: torch.Tensor, k_scale: torch.Tensor, attn_mask=None,
                  is_causal: bool = False, sm_scale: float = 1.0, out_dtype: torch.dtype = torch.bfloat16, return_lse: bool = False,
                  v_scale: torch.Tensor | None = None):
    """K9: sdnq_atten_fwd (kernels/triton_atten.py:338-386) for 1-byte q / k codes with per-row scales and a 16-bit v.
    q_q [Z,H,QN,HD], k_q [Z,KH,KN,HD] int8 / float8_e4m3fn; q_scale [Z,H,QN], k_scale [Z,KH,KN] f32; v [Z,VH,KN,HDV] bf16 / f16;
    attn_mask: None, or a 4-D int8 / bool (0 = masked out) or float (additive) tensor broadcastable to [Z,H,QN,KN].
    -> (out [Z,H,QN,HDV], lse [Z,H,QN] | None)."""
    import ctypes
    _require_cuda(q_q, k_q, v, q_scale, k_scale)
    if q_q.dtype != k_q.dtype or q_q.dtype not in (torch.int8, torch.float8_e4m3fn):
        raise _lib.SDNQKernelError(f"attention: q / k codes must both be int8 or float8_e4m3fn (got {q_q.dtype}, {k_q.dtype})")
    v_is_codes = v.dtype in (torch.int8, torch.float8_e4m3fn)
    if v_is_codes != (v_scale is not None):
        raise _lib.SDNQKernelError(f"attention: v_scale goes with int8 / float8_e4m3fn v codes, and only with them (v is {v.dtype})")
    if v_scale is not None:
        _require_cuda(v_scale)
        v_scale = v_scale.to(torch.float32).contiguous()
        if tuple(v_scale.shape) != tuple(v.shape[:-1]):
            raise _lib.SDNQKernelError(f"attention: v_scale {tuple(v_scale.shape)} does not match v {tuple(v.shape)}")
    q_q, k_q, v = q_q.contiguous(), k_q.contiguous(), v.contiguous()
    q_scale = q_scale.to(torch.float32).contiguous()
    k_scale = k_scale.to(torch.float32).contiguous()
    Z, H, QN, HD = q_q.shape
    _, KH, KN, _ = k_q.shape
    _, VH, VN, HDV = v.shape
    if VN != KN or k_q.shape[0] != Z or v.shape[0] != Z or k_q.shape[3] != HD:
        raise _lib.SDNQKernelError(f"attention: inconsistent shapes q {tuple(q_q.shape)} k {tuple(k_q.shape)} v {tuple(v.shape)}")
    dev = q_q.device
    mask_ptr, mask_code, strides = None, 0, None
    if attn_mask is not None:
        mk = attn_mask
        if mk.dtype == torch.bool:
            mk = mk.to(torch.int8)
        elif mk.dtype != torch.int8:
            mk = mk.to(torch.float32)
        while mk.ndim < 4:
            mk = mk.unsqueeze(0)
        if mk.shape[-1] == 1 and KN != 1:
            mk = mk.expand(-1, -1, -1, KN)
        attn_mask = mk = mk.contiguous()                                                   # kept alive until the launch below
        st = [mk.stride(i) if mk.shape[i] != 1 else 0 for i in range(4)]                  # triton_atten.py:370-376
        strides = (ctypes.c_int64 * 4)(*st)
        mask_ptr, mask_code = mk.data_ptr(), dtype_code(mk.dtype) if mk.dtype != torch.int8 else SDNQ_I8
    out = torch.empty((Z, H, QN, HDV), dtype=out_dtype, device=dev)
    lse = torch.empty((Z, H, QN), dtype=out_dtype, device=dev) if return_lse else None
    lib = _lib.load()
    ws_bytes = int(lib.sdnq_b200_attention_workspace_bytes(Z, VH, KN, HDV))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.sdnq_b200_attention(_ptr(q_q), _ptr(k_q), _ptr(v), SDNQ_I8 if q_q.dtype == torch.int8 else SDNQ_F8E4M3, dtype_code(v.dtype),
                                      _ptr(q_scale), _ptr(k_scale), _ptr(v_scale), mask_ptr, mask_code, strides, _ptr(out), _ptr(lse), dtype_code(out_dtype),
                                      Z, H, KH, VH, QN, KN, HD, HDV, float(sm_scale), int(bool(is_causal)), _ptr(ws), ws_bytes, _stream(q_q)))
    return out, lse


class DequantBatch:
    """A planned batched dequantisation: the device-side table of sdnq_b200_dequant_batch_plan, the [N,K] outputs it writes (views of
    `slab`) and everything the embedded pointers refer to."""
    __slots__ = ("table", "info", "outs", "keep", "device")

    def __init__(self, table, info, outs, keep, device):
        self.table, self.info, self.outs, self.keep, self.device = table, info, outs, keep, device


def dequant_batch_bytes(shapes, out_dtype=torch.bfloat16) -> list:
    """byte offsets of the [N,K] outputs of a batch inside one slab (256-byte aligned), plus the total"""
    offs, pos = [], 0
    esz = torch.empty((), dtype=out_dtype).element_size()
    for n, k in shapes:
        offs.append(pos)
        pos += (n * k * esz + 255) // 256 * 256
    return offs + [pos]


def dequant_batch_plan(jobs, slab: torch.Tensor, out_dtype=torch.bfloat16) -> DequantBatch:
    """jobs: dicts with the arguments of `dequant` (weight, weights_dtype, scale, zero_point, N, K, group_size, svd_up, svd_down,
    svd_layout_matmul).  The outputs are laid out in `slab` (uint8, device) at the offsets of dequant_batch_bytes.  Synchronous
    (uploads the plan): not to be called while the stream is being captured.  Raises SDNQKernelError if a job is not covered."""
    import ctypes
    lib = _lib.load()
    n = len(jobs)
    offs = dequant_batch_bytes([(j["N"], j["K"]) for j in jobs], out_dtype)
    if offs[-1] > slab.numel() or slab.dtype != torch.uint8 or slab.data_ptr() % 16 != 0:
        raise _lib.SDNQKernelError("dequant_batch_plan: the slab is too small or misaligned")
    arr = (_lib.DequantJob * n)()
    keep, outs = [], []
    for i, j in enumerate(jobs):
        e = dtype_dict[j["weights_dtype"]]
        w = j["weight"].contiguous() if e["is_packed"] else physical_nk(j["weight"])
        scale = j["scale"]
        scale = scale.to(torch.float32).contiguous() if scale.dtype != torch.float32 or not scale.is_contiguous() else scale
        zp = j.get("zero_point")
        if zp is not None and (zp.dtype != torch.float32 or not zp.is_contiguous()):
            zp = zp.to(torch.float32).contiguous()
        up, down = j["svd_up"], j["svd_down"]
        _require_cuda(w, scale, up, down)
        N, K = int(j["N"]), int(j["K"])
        out = slab[offs[i]:offs[i] + N * K * 2].view(out_dtype).view(N, K)
        q = arr[i]
        q.weight, q.fmt, q.scale, q.zero_point = w.data_ptr(), weight_format(j["weights_dtype"], w), scale.data_ptr(), _ptr(zp)
        q.N, q.K, q.group_size = N, K, _group_args(scale, N, K, j["group_size"], False, e["num_bits"])
        if j.get("svd_layout_matmul"):      # svd_up [r,N], svd_down [K,r]
            q.svd_rank = up.shape[0]
            q.svd_up, q.up_stride_n, q.up_stride_r = up.data_ptr(), up.stride(1), up.stride(0)
            q.svd_down, q.down_stride_r, q.down_stride_k = down.data_ptr(), down.stride(1), down.stride(0)
        else:                               # svd_up [N,r], svd_down [r,K]
            q.svd_rank = up.shape[1]
            q.svd_up, q.up_stride_n, q.up_stride_r = up.data_ptr(), up.stride(0), up.stride(1)
            q.svd_down, q.down_stride_r, q.down_stride_k = down.data_ptr(), down.stride(0), down.stride(1)
        q.svd_dtype, q.out, q.out_dtype = dtype_code(up.dtype), out.data_ptr(), dtype_code(out_dtype)
        # strong references only to what was created here (converted copies): the layer's own tensors are guarded by the caller's
        # identity check (forward._StoredState) before every run, and a plan must not keep a deleted model's weights alive
        keep += [t for t, src in ((w, j["weight"]), (scale, j["scale"]), (zp, j.get("zero_point"))) if t is not None and t is not src]
        outs.append(out)
    nbytes = int(lib.sdnq_b200_dequant_batch_table_bytes(n))
    host = torch.empty(nbytes + 128, dtype=torch.uint8)
    hoff = (-host.data_ptr()) % 128
    info = (ctypes.c_int32 * 4)()
    check(lib.sdnq_b200_dequant_batch_plan(ctypes.addressof(arr), n, host.data_ptr() + hoff, ctypes.addressof(info)))
    table = torch.empty(nbytes + 128, dtype=torch.uint8, device=slab.device)
    doff = (-table.data_ptr()) % 128
    table = table[doff:doff + nbytes]
    table.copy_(host[hoff:hoff + nbytes])
    torch.cuda.current_stream(slab.device).synchronize()
    return DequantBatch(table, info, outs, keep + [slab], slab.device)


def dequant_batch_run(plan: DequantBatch):
    """one launch: every weight of the plan dequantised into its output (current stream of the plan's device)"""
    import ctypes
    with torch.cuda.device(plan.device):
        check(_lib.load().sdnq_b200_dequant_batch_run(plan.table.data_ptr(), ctypes.addressof(plan.info), torch.cuda.current_stream(plan.device).cuda_stream))


def dequant_nd(weight, weights_dtype, scale, zero_point, view_shape, out_dtype, use_codebook=False, addend=None) -> torch.Tensor:
    """K3 for convolution layers: broadcast dequant over the quantised view (scale / zero_point broadcastable to view_shape;
    codebook levels carry one extra trailing axis).  `addend` (same numel) is added in f32 before the cast."""
    import ctypes
    _require_cuda(weight, scale)
    e = dtype_dict[weights_dtype]
    w = weight.contiguous()
    view_shape = tuple(int(v) for v in view_shape)
    scale = scale.to(torch.float32) if scale.dtype != torch.float32 else scale
    sshape = tuple(scale.shape[:-1]) if use_codebook else tuple(scale.shape)
    if len(sshape) != len(view_shape):
        raise _lib.SDNQKernelError(f"dequant_nd: scale shape {tuple(scale.shape)} does not line up with the quantised view {view_shape}")
    scale = scale.contiguous()
    strides, acc = [], 1
    for dim, sdim in zip(reversed(view_shape), reversed(sshape)):
        if sdim == 1:
            strides.append(0)
        elif sdim == dim:
            strides.append(acc)
            acc *= sdim
        else:
            raise _lib.SDNQKernelError(f"dequant_nd: scale shape {tuple(scale.shape)} is not broadcastable to {view_shape}")
    strides.reverse()
    if zero_point is not None:
        zero_point = zero_point.to(torch.float32).contiguous()
        if tuple(zero_point.shape) != tuple(scale.shape):
            raise _lib.SDNQKernelError("dequant_nd: zero_point and scale shapes differ")
    out = torch.empty(view_shape, dtype=out_dtype, device=w.device)
    n = len(view_shape)
    dims_c = (ctypes.c_int64 * n)(*view_shape)
    str_c = (ctypes.c_int64 * n)(*strides)
    if addend is not None:
        addend = addend.contiguous()
        if addend.numel() != out.numel():
            raise _lib.SDNQKernelError("dequant_nd: addend size mismatch")
    fmt = weight_format(weights_dtype, w)
    with torch.cuda.device(w.device):
        check(_lib.load().sdnq_b200_dequant_nd(_ptr(w), fmt, _ptr(scale), _ptr(zero_point), int(use_codebook), n, dims_c, str_c, _ptr(addend),
                                               dtype_code(addend.dtype) if addend is not None else SDNQ_F32, _ptr(out), dtype_code(out_dtype), _stream(w)))
    return out


def quantize_weight(w: torch.Tensor, weights_dtype: str, group_size: int = -1, scale_dtype: torch.dtype | None = None):
    """K8.  w [N,K] f32 / bf16 / f16 -> (codes, scale [N, K/g] f32, zero_point [N, K/g] f32 | None): `codes` is the packed uint8
    buffer (sub-byte formats) or the [N,K] code matrix (8-bit formats) exactly as quantize_weight + pack_int / pack_float produce them
    (integer formats, torch.float8_e4m3fn / float8_e5m2, and the eXmY minifloats of packed_float.py)."""
    _require_cuda(w)
    info = dtype_dict[weights_dtype]
    if not 2 <= info["num_bits"] <= 8 or not (info["is_integer"] or info["is_packed"] or info["torch_dtype"] in (torch.float8_e4m3fn, torch.float8_e5m2)):
        raise _lib.SDNQKernelError(f"quantize_weight: {weights_dtype!r} has no quantisation kernel (integer and float formats of 2..8 bits)")
    N, K = w.shape
    w = w.contiguous()
    if w.data_ptr() % 16 != 0:
        w = w.clone()
    g = K if group_size <= 0 or group_size >= K else int(group_size)
    bits = info["num_bits"]
    # 8-bit formats: the [N,K] code matrix in the storage dtype (int8 / uint8 / float8_*; uint8 for the 8-bit eXmY minifloats)
    codes = torch.empty((N, K), dtype=info["storage_dtype"], device=w.device) if bits == 8 else torch.empty(N * K * bits // 8, dtype=torch.uint8, device=w.device)
    scale = torch.empty((N, K // g), dtype=torch.float32, device=w.device)
    zp = torch.empty((N, K // g), dtype=torch.float32, device=w.device) if info["is_unsigned"] else None
    with torch.cuda.device(w.device):
        check(_lib.load().sdnq_b200_quantize_weight(_ptr(w), dtype_code(w.dtype), N, K, g, weight_format(weights_dtype), dtype_code(scale_dtype or torch.float32),
                                                    _ptr(codes), _ptr(scale), _ptr(zp), _stream(w)))
    return codes, scale, zp


def requant(weight, weights_dtype, scale, zero_point, N, K, group_size, matmul_dtype, use_codebook=False, want_colsum=False):
    """K4.  Returns (wq [N,K] physical, sw [N], zw [N] | None, colsum [N] | None)."""
    _require_cuda(weight, scale)
    lib = _lib.load()
    e = dtype_dict[weights_dtype]
    w = weight.contiguous() if e["is_packed"] else physical_nk(weight)
    scale = scale.to(torch.float32).contiguous()
    if zero_point is not None:
        zero_point = zero_point.to(torch.float32).contiguous()
    code = mm_code(matmul_dtype)
    fmt = weight_format(weights_dtype, w)
    gs = _group_args(scale, N, K, group_size, use_codebook, e["num_bits"])
    wq = torch.empty((N, K), dtype=_MM_TORCH[code], device=w.device)
    sw = torch.empty((N,), dtype=torch.float32, device=w.device)
    zw = torch.empty((N,), dtype=torch.float32, device=w.device) if code == SDNQ_U8 else None
    colsum = torch.empty((N,), dtype=torch.int32, device=w.device) if (want_colsum or code == SDNQ_U8) else None
    with torch.cuda.device(w.device):
        check(lib.sdnq_b200_requant(_ptr(w), fmt, _ptr(scale), _ptr(zero_point), int(use_codebook), N, K, gs, code,
                                    _ptr(wq), _ptr(sw), _ptr(zw), _ptr(colsum), _stream(w)))
    return wq, sw, zw, colsum


def act_quant(x: torch.Tensor, matmul_dtype: str, hadamard_group: int = 0, want_rowsum: bool = False, want_x_rot: bool = False):
    """K2.  x [..., K] -> (xq [M,K], sx [M], zx [M] | None, rowsum [M] | None, x_rot [M,K] | None)."""
    _require_cuda(x)
    lib = _lib.load()
    K = x.shape[-1]
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1 or x2.stride(0) % 8 != 0 or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M = x2.shape[0]
    code = mm_code(matmul_dtype)
    dev = x2.device
    xq = torch.empty((M, K), dtype=_MM_TORCH[code], device=dev)
    sx = torch.empty((M,), dtype=torch.float32, device=dev)
    zx = torch.empty((M,), dtype=torch.float32, device=dev) if code == SDNQ_U8 else None
    rowsum = torch.empty((M,), dtype=torch.int32, device=dev) if want_rowsum else None
    x_rot = torch.empty((M, K), dtype=x2.dtype, device=dev) if want_x_rot else None
    with torch.cuda.device(dev):
        check(lib.sdnq_b200_act_quant(_ptr(x2), dtype_code(x2.dtype), M, K, x2.stride(0), int(hadamard_group), code,
                                      _ptr(xq), _ptr(sx), _ptr(zx), _ptr(rowsum), _ptr(x_rot), _stream(x2)))
    return xq, sx, zx, rowsum, x_rot


def conv_act_quant(x: torch.Tensor, kernel_size, stride, padding, dilation, matmul_dtype: str, hadamard_group: int = 0,
                   want_rowsum: bool = False, want_x_rot: bool = False):
    """K2 over the im2col view of a conv input (never materialised).  x [B,C,H,W] (any non-negative strides; conv1d callers pass
    H = 1) -> (xq [M,K], sx [M], zx, rowsum, x_rot, (B, H_out, W_out)) with M = B*H_out*W_out, K = C*kh*kw in (c, i, j) order."""
    _require_cuda(x)
    if x.ndim != 4:
        raise _lib.SDNQKernelError("conv_act_quant expects a 4-D input (conv1d: unsqueeze(2))")
    if any(s < 0 for s in x.stride()):
        x = x.contiguous()
    B, C, H, W = x.shape
    (kh, kw), (sh, sw), (ph, pw), (dh, dw) = kernel_size, stride, padding, dilation
    Hout = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    Wout = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    M, K = B * Hout * Wout, C * kh * kw
    code = mm_code(matmul_dtype)
    dev = x.device
    xq = torch.empty((M, K), dtype=_MM_TORCH[code], device=dev)
    sx = torch.empty((M,), dtype=torch.float32, device=dev)
    zx = torch.empty((M,), dtype=torch.float32, device=dev) if code == SDNQ_U8 else None
    rowsum = torch.empty((M,), dtype=torch.int32, device=dev) if want_rowsum else None
    x_rot = torch.empty((M, K), dtype=x.dtype, device=dev) if want_x_rot else None
    geo = _lib.Conv2dGeometry(B, C, H, W, *x.stride(), kh, kw, sh, sw, ph, pw, dh, dw)
    lib = _lib.load()
    # scratch for the tiled path (per-input-pixel channel statistics); the library falls back to the gather kernel where the
    # tiled one does not apply (other kernel sizes, rotation, x_rot)
    ws = torch.empty(max(int(lib.sdnq_b200_conv_act_quant_workspace_bytes(geo, code)), 8), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.sdnq_b200_conv_act_quant_ws(_ptr(x), dtype_code(x.dtype), geo, int(hadamard_group), code, _ptr(xq), _ptr(sx), _ptr(zx),
                                              _ptr(rowsum), _ptr(x_rot), _ptr(ws), ws.numel(), _stream(x)))
    return xq, sx, zx, rowsum, x_rot, (B, Hout, Wout)


def rows_to_nchw(y: torch.Tensor, batch: int, hw: int) -> torch.Tensor:
    """[batch*hw, N] GEMM output -> contiguous [batch, N, hw] (the permute(0, 3, 1, 2).contiguous() of the conv forwards)."""
    _require_cuda(y)
    y = y.contiguous()
    N = y.shape[-1]
    out = torch.empty((batch, N, hw), dtype=y.dtype, device=y.device)
    if y.element_size() not in (2, 4):
        raise _lib.SDNQKernelError(f"rows_to_nchw: unsupported dtype {y.dtype}")
    with torch.cuda.device(y.device):
        check(_lib.load().sdnq_b200_rows_to_nchw(_ptr(y), _ptr(out), y.element_size(), batch, hw, N, _stream(y)))
    return out


def scaled_mm(a: torch.Tensor, b_nk: torch.Tensor, sx: torch.Tensor, sw: torch.Tensor, bias: torch.Tensor | None = None,
              out_dtype: torch.dtype = torch.bfloat16, rowsum=None, zp=None, colsum=None, zx=None) -> torch.Tensor:
    """K1.  a [M,K] int8/fp8 contiguous, b_nk the physical [N,K] weight.  bias None | [N] | [M,N]."""
    _require_cuda(a, b_nk)
    lib = _lib.load()
    M, K = a.shape
    N = b_nk.shape[0]
    assert b_nk.shape[1] == K and a.is_contiguous() and b_nk.is_contiguous()
    out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    bias_ld, bias_code = 0, SDNQ_F32
    if bias is not None:
        bias = bias.contiguous()
        bias_code = dtype_code(bias.dtype)
        if bias.ndim == 2 and bias.shape[0] != 1:
            bias_ld = bias.stride(0)
    ab = _operand_code(a.dtype, b_nk)
    stream = _stream(a)
    ws = _gemm_workspace(a.device, stream) if ab == SDNQ_I8 else None      # stream-K scratch (int8 only)
    with torch.cuda.device(a.device):
        check(lib.sdnq_b200_scaled_mm_ws(_ptr(a), _ptr(b_nk), ab, _ptr(sx), _ptr(sw), _ptr(bias), bias_code, bias_ld,
                                         _ptr(rowsum), _ptr(zp), _ptr(colsum), _ptr(zx), _ptr(out), dtype_code(out_dtype), M, N, K,
                                         _ptr(ws), 0 if ws is None else ws.numel(), stream))
    return out


_GEMM_WS: dict = {}


def _gemm_workspace(device, stream: int):
    """Per (device, stream, graph capture) scratch of the stream-K schedule of K1 (parked partial accumulators + flags, zeroed once: the
    kernel leaves the flags zero after every launch).  One launch at a time uses it: launches on one stream are ordered."""
    key = (device.index, stream, capture_id(stream))      # a graph capture gets its own (allocated from the graph's pool, zeroed by a captured memset)
    ws = _GEMM_WS.get(key)
    if ws is None:
        if len(_GEMM_WS) > 64:
            _GEMM_WS.clear()
        ws = _GEMM_WS[key] = torch.zeros(int(_lib.load().sdnq_b200_scaled_mm_workspace_bytes()), dtype=torch.uint8, device=device)
    return ws


def scaled_mm_packed(a: torch.Tensor, b_packed: torch.Tensor, weights_dtype: str, N: int, sx, sw, bias=None,
                     out_dtype: torch.dtype = torch.bfloat16, rowsum=None, zp=None) -> torch.Tensor:
    """K1 with in-kernel unpack: a [M,K] int8, b_packed the stored row-wise packed int4 / uint4 weight ([N*K/2] bytes)."""
    _require_cuda(a, b_packed)
    M, K = a.shape
    out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    bias_ld, bias_code = 0, SDNQ_F32
    if bias is not None:
        bias = bias.contiguous()
        bias_code = dtype_code(bias.dtype)
        if bias.ndim == 2 and bias.shape[0] != 1:
            bias_ld = bias.stride(0)
    fmt = weight_format(weights_dtype, b_packed)
    with torch.cuda.device(a.device):
        check(_lib.load().sdnq_b200_scaled_mm_packed(_ptr(a), _ptr(b_packed), fmt, _ptr(sx), _ptr(sw), _ptr(bias), bias_code, bias_ld,
                                                     _ptr(rowsum), _ptr(zp), _ptr(out), dtype_code(out_dtype), M, N, K, _stream(a)))
    return out


def svd_low(x: torch.Tensor, svd_down_rk: torch.Tensor) -> torch.Tensor:
    """K7.  x [M,K] bf16 / f16 (row stride a multiple of 8), svd_down_rk [r,K] contiguous of the same dtype -> low [M,r] = cast(x @ down^T)."""
    _require_cuda(x, svd_down_rk)
    M, K = x.shape
    r = svd_down_rk.shape[0]
    if svd_down_rk.dtype != x.dtype or not svd_down_rk.is_contiguous() or svd_down_rk.shape[1] != K:
        raise _lib.SDNQKernelError("svd_low: svd_down must be a contiguous [r,K] tensor of the activation dtype")
    if x.stride(-1) != 1 or x.stride(0) % 8 != 0 or x.data_ptr() % 16 != 0:
        x = x.contiguous()
    low = torch.empty((M, r), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().sdnq_b200_svd_low(_ptr(x), dtype_code(x.dtype), x.stride(0), _ptr(svd_down_rk), r, _ptr(low), M, K, _stream(x)))
    return low


def scaled_mm_svd(a: torch.Tensor, b: torch.Tensor, sx, sw, low: torch.Tensor, svd_up_nr: torch.Tensor, bias=None,
                  out_dtype: torch.dtype = torch.bfloat16, rowsum=None, zp=None, colsum=None, zx=None,
                  packed_dtype: str | None = None, N: int | None = None) -> torch.Tensor:
    """K1 with the SVD rank-r term accumulated on the tensor cores: out = scaled_mm(...) + low @ svd_up_nr^T (added to the bias in f32).
    b is the physical [N,K] operand, or (packed_dtype = 'int4' / 'uint4', N given) the stored packed weight."""
    _require_cuda(a, b, low, svd_up_nr)
    M, K = a.shape
    N = b.shape[0] if packed_dtype is None else int(N)
    r = svd_up_nr.shape[1]
    if (low.dtype != svd_up_nr.dtype or tuple(low.shape) != (M, r) or tuple(svd_up_nr.shape) != (N, r)
            or not low.is_contiguous() or not svd_up_nr.is_contiguous()):
        raise _lib.SDNQKernelError("scaled_mm_svd: low [M,r] / svd_up [N,r] must be contiguous tensors of one 16-bit dtype")
    out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    bias_ld, bias_code = 0, SDNQ_F32
    if bias is not None:
        bias = bias.contiguous()
        bias_code = dtype_code(bias.dtype)
        if bias.ndim == 2 and bias.shape[0] != 1:
            bias_ld = bias.stride(0)
    if packed_dtype is None:
        assert b.shape[1] == K and a.is_contiguous() and b.is_contiguous()
        ab, fmt = _operand_code(a.dtype, b), None
    else:
        ab, fmt = SDNQ_I8, weight_format(packed_dtype, b)
    with torch.cuda.device(a.device):
        check(_lib.load().sdnq_b200_scaled_mm_svd(_ptr(a), _ptr(b), ab, fmt, _ptr(sx), _ptr(sw), _ptr(bias), bias_code, bias_ld,
                                                  _ptr(rowsum), _ptr(zp), _ptr(colsum), _ptr(zx), _ptr(low), _ptr(svd_up_nr), r,
                                                  dtype_code(low.dtype), _ptr(out), dtype_code(out_dtype), M, N, K, _stream(a)))
    return out


def scaled_mm_grouped(a: torch.Tensor, b_cat: torch.Tensor, sx, sw_cat, starts, ns, bias_cat=None, out_dtype: torch.dtype = torch.bfloat16,
                      rowsum=None, zp=None, colsum=None, zx=None, packed_dtype: str | None = None) -> list:
    """K1 over sibling projections in one launch.  a [M,K] codes; b_cat the siblings' [N_g,K] operands (or packed int4 / uint4 rows)
    stacked, segment g at rows starts[g] .. starts[g] + ns[g] (starts has len(ns) + 1 entries, multiples of 128); per-channel
    vectors laid out the same way.  Returns one contiguous [M, ns[g]] tensor per sibling."""
    import ctypes
    _require_cuda(a, b_cat)
    M, K = a.shape
    G = len(ns)
    assert a.is_contiguous() and b_cat.is_contiguous() and len(starts) == G + 1 and b_cat.shape[0] == starts[-1]
    outs = [torch.empty((M, n), dtype=out_dtype, device=a.device) for n in ns]
    if packed_dtype is None:
        assert b_cat.shape[1] == K
        ab, fmt = _operand_code(a.dtype, b_cat), None
    else:
        ab, fmt = (SDNQ_I8 if a.dtype == torch.int8 else SDNQ_F8E4M3), weight_format(packed_dtype, b_cat)
    bias_code = dtype_code(bias_cat.dtype) if bias_cat is not None else SDNQ_F32
    starts_c = (ctypes.c_int64 * (G + 1))(*[int(v) for v in starts])
    ns_c = (ctypes.c_int64 * G)(*[int(v) for v in ns])
    outs_c = (ctypes.c_void_p * G)(*[o.data_ptr() for o in outs])
    with torch.cuda.device(a.device):
        check(_lib.load().sdnq_b200_scaled_mm_grouped(_ptr(a), _ptr(b_cat), ab, fmt, _ptr(sx), _ptr(sw_cat), _ptr(bias_cat), bias_code,
                                                      _ptr(rowsum), _ptr(zp), _ptr(colsum), _ptr(zx), G, starts_c, ns_c, outs_c,
                                                      dtype_code(out_dtype), M, K, _stream(a)))
    return outs


def mm(a: torch.Tensor, b_nk: torch.Tensor) -> torch.Tensor:
    """plain int8 -> int32 / fp8 -> f32 matmul (int_mm_func / fp8_mm_func)."""
    _require_cuda(a, b_nk)
    M, K = a.shape
    N = b_nk.shape[0]
    fp8 = a.dtype == torch.float8_e4m3fn
    out = torch.empty((M, N), dtype=torch.float32 if fp8 else torch.int32, device=a.device)
    with torch.cuda.device(a.device):
        check(_lib.load().sdnq_b200_mm(_ptr(a), _ptr(b_nk), _operand_code(a.dtype, b_nk), _ptr(out), M, N, K, _stream(a)))
    return out


def linear_small_m(x: torch.Tensor, wq_nk: torch.Tensor, sw, zp=None, bias=None) -> torch.Tensor:
    """K5: x [..., K] with fewer than 33 rows times the 1-byte weight codes wq_nk [N,K] (int8 / float8_e4m3fn, row-wise scales)."""
    _require_cuda(x, wq_nk)
    K = x.shape[-1]
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1 or x2.stride(0) % 8 != 0 or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M, N = x2.shape[0], wq_nk.shape[0]
    out = torch.empty((M, N), dtype=x.dtype, device=x.device)
    if wq_nk.dtype not in (torch.int8, torch.float8_e4m3fn, torch.float8_e5m2):
        raise _lib.SDNQKernelError(f"linear_small_m: weight codes must be int8 / float8_e4m3fn / float8_e5m2 (got {wq_nk.dtype})")
    wcode = dtype_code(wq_nk.dtype)
    with torch.cuda.device(x.device):
        check(_lib.load().sdnq_b200_linear_small_m(_ptr(x2), dtype_code(x2.dtype), x2.stride(0), _ptr(wq_nk), wcode, _ptr(sw), _ptr(zp), _ptr(bias),
                                                   dtype_code(bias.dtype) if bias is not None else SDNQ_F32, _ptr(out), M, N, K, _stream(x)))
    return out.view(*x.shape[:-1], N)


def linear_small_m_packed(x: torch.Tensor, weight: torch.Tensor, weights_dtype: str, scale: torch.Tensor, zero_point, N: int, K: int,
                          bias: torch.Tensor | None = None) -> torch.Tensor:
    """K5p: x [..., K] with fewer than 33 rows times the *stored* weight (packed sub-byte / minifloat / 8-bit codes of the flattened
    [N,K] tensor, scale / zero_point [N, K/g] in any shape with that many elements).  bias None | [N] | [M,N]."""
    _require_cuda(x, weight, scale)
    e = dtype_dict[weights_dtype]
    w = weight.contiguous() if e["is_packed"] else physical_nk(weight)
    scale = scale.to(torch.float32).contiguous()
    if zero_point is not None:
        zero_point = zero_point.to(torch.float32).contiguous()
    per_row = scale.numel() // N
    if per_row < 1 or scale.numel() != N * per_row or K % per_row != 0:
        raise _lib.SDNQKernelError(f"linear_small_m_packed: {scale.numel()} scales do not tile a [{N},{K}] weight")
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1 or x2.stride(0) % 8 != 0 or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M = x2.shape[0]
    out = torch.empty((M, N), dtype=x.dtype, device=x.device)
    bias_ld, bias_code = 0, SDNQ_F32
    if bias is not None:
        bias = bias.contiguous()
        bias_code = dtype_code(bias.dtype)
        if bias.ndim == 2 and bias.shape[0] != 1:
            bias_ld = bias.stride(0)
    fmt = weight_format(weights_dtype, w)
    with torch.cuda.device(x.device):
        check(_lib.load().sdnq_b200_linear_small_m_packed(_ptr(x2), dtype_code(x2.dtype), x2.stride(0), _ptr(w), fmt, _ptr(scale), _ptr(zero_point),
                                                          K // per_row, _ptr(bias), bias_code, bias_ld, _ptr(out), M, N, K, _stream(x)))
    return out.view(*x.shape[:-1], N)


def linear_w4a16(x: torch.Tensor, weight: torch.Tensor, weights_dtype: str, scale: torch.Tensor, zero_point, N: int, K: int,
                 bias: torch.Tensor | None = None, svd_down_rk: torch.Tensor | None = None, svd_up_nr: torch.Tensor | None = None) -> torch.Tensor:
    """K6: the dequant-path Linear in one launch.  x [..., K] bf16 / f16; `weight` the stored packed int4 / uint4 tensor of the
    [N,K] weight; scale / zero_point with N * (K / group) elements; svd factors already as [r,K] / [N,r] contiguous in x.dtype."""
    _require_cuda(x, weight, scale)
    w = weight.contiguous()
    scale = scale.to(torch.float32).contiguous()
    if zero_point is not None:
        zero_point = zero_point.to(torch.float32).contiguous()
    per_row = scale.numel() // N
    if per_row < 1 or scale.numel() != N * per_row or K % per_row != 0:
        raise _lib.SDNQKernelError(f"linear_w4a16: {scale.numel()} scales do not tile a [{N},{K}] weight")
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1 or x2.stride(0) % 8 != 0 or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M = x2.shape[0]
    out = torch.empty((M, N), dtype=x.dtype, device=x.device)
    rank = 0
    if svd_up_nr is not None:
        rank = svd_up_nr.shape[1]
        if (svd_up_nr.dtype != x.dtype or svd_down_rk.dtype != x.dtype or not svd_up_nr.is_contiguous() or not svd_down_rk.is_contiguous()
                or tuple(svd_up_nr.shape) != (N, rank) or tuple(svd_down_rk.shape) != (rank, K)):
            raise _lib.SDNQKernelError("linear_w4a16: svd factors must be contiguous [N,r] / [r,K] tensors of the activation dtype")
    if bias is not None:
        bias = bias.contiguous()
    fmt = weight_format(weights_dtype, w)
    with torch.cuda.device(x.device):
        check(_lib.load().sdnq_b200_linear_w4a16(_ptr(x2), dtype_code(x2.dtype), x2.stride(0), _ptr(w), fmt, _ptr(scale), _ptr(zero_point), K // per_row,
                                                 _ptr(svd_down_rk), _ptr(svd_up_nr), rank, _ptr(bias), dtype_code(bias.dtype) if bias is not None else SDNQ_F32,
                                                 _ptr(out), M, N, K, _stream(x)))
    return out.view(*x.shape[:-1], N)


_WORKSPACES: dict = {}
_RETIRED_WORKSPACES: list = []


def _workspace(dev, nbytes):
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            # a captured CUDA graph may still hold the old pointer: outgrown workspaces are kept alive, never freed
            # (sizes grow geometrically below, so this is a handful of buffers per stream at most)
            _RETIRED_WORKSPACES.append(ws)
            nbytes = max(nbytes, 2 * ws.numel())
        # zero-filled: the first 4 KB hold the fused kernel's cross-CTA strip counters, which every launch leaves at zero
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _WORKSPACES[key] = ws
    return ws


def linear_w8a8(x, wq_nk, matmul_dtype, sw, bias=None, zp=None, colsum=None, hadamard_group=0, out_dtype=None, fused=None) -> torch.Tensor:
    """The W8A8 Linear in one C call (per-stream cached workspace): K2 + K1 chained by programmatic dependent launch.
    fused=True asks for the single-launch kernel (in-kernel activation quantiser; raises when the case is not covered),
    fused=None leaves the choice to the library (SDNQ_B200_FUSED)."""
    _require_cuda(x, wq_nk)
    lib = _lib.load()
    K = x.shape[-1]
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1 or x2.stride(0) % 8 != 0 or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M, N = x2.shape[0], wq_nk.shape[0]
    out_dtype = out_dtype or x.dtype
    out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    nbytes = lib.sdnq_b200_linear_w8a8_workspace_bytes(M, K)
    ws = _workspace(x.device, nbytes)
    code = mm_code(matmul_dtype)
    expect = torch.int8 if code in (SDNQ_I8, SDNQ_U8) else torch.float8_e4m3fn
    if wq_nk.dtype == torch.float8_e5m2 and code == SDNQ_F8E4M3:
        code = SDNQ_F8E5M2                    # e4m3 activation codes x the stored e5m2 weight
    elif wq_nk.dtype != expect:
        raise _lib.SDNQKernelError(f"linear_w8a8: a {matmul_dtype} matmul cannot read {wq_nk.dtype} weight codes")
    if fused:
        if zp is not None or colsum is not None or hadamard_group:
            raise _lib.SDNQKernelError("fused Linear: no zero-point terms and no Hadamard rotation")
        with torch.cuda.device(x.device):
            check(lib.sdnq_b200_linear_w8a8_fused(_ptr(x2), dtype_code(x2.dtype), x2.stride(0), _ptr(wq_nk), code, _ptr(sw),
                                                  _ptr(bias), dtype_code(bias.dtype) if bias is not None else SDNQ_F32, _ptr(out),
                                                  dtype_code(out_dtype), M, N, K, _ptr(ws), ws.numel(), _stream(x)))
        return out.view(*x.shape[:-1], N)
    with torch.cuda.device(x.device):
        check(lib.sdnq_b200_linear_w8a8(_ptr(x2), dtype_code(x2.dtype), x2.stride(0), _ptr(wq_nk), code, _ptr(sw),
                                        _ptr(zp), _ptr(colsum), _ptr(bias), dtype_code(bias.dtype) if bias is not None else SDNQ_F32,
                                        int(hadamard_group), _ptr(out), dtype_code(out_dtype), M, N, K, _ptr(ws), ws.numel(), _stream(x)))
    return out.view(*x.shape[:-1], N)
