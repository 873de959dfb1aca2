"""The hot path: forward functions plugged into SDNQLayer.forward_func.

Same names / dispatch rule as the reference (forward.py:6-57, layers/linear/*.py) so `get_forward_func(...)` is a drop-in:

    quantized_linear_forward                  K3 dequant kernel -> bf16 GEMM (library)          layers/linear/forward.py:24-26
    quantized_linear_forward_int8_matmul      K2 act-quant + K1 tcgen05 int8 GEMM               layers/linear/linear_int8.py:100-125
    quantized_linear_forward_uint8_matmul     K2 (asymmetric) + K1 with zero-point epilogue     layers/linear/linear_uint8.py:105-129
    quantized_linear_forward_fp8_matmul       K2 (e4m3) + K1 tcgen05 fp8 GEMM                   layers/linear/linear_fp8.py:81-104
    quantized_conv_forward (+ transpose 1/2/3d) K3 broadcast dequant -> library convolution       layers/conv/forward.py:79-99
    quantized_conv_forward_{int8,uint8,fp8}_matmul  K2 over the im2col view (gathered, never materialised) + K1   layers/conv/conv_*.py

Every function runs CUDA kernels from libsdnq_b200.so; there is no eager / CPU fallback (a CPU tensor raises)."""
import os
import warnings
import weakref
from collections.abc import Callable

import torch

from . import ops, prefetch, siblings
from .common import conv_transpose_types, conv_types, dtype_dict, embedding_types

SMALL_M = 32     # rows below this use the dequant path, as upstream does (linear_int8.py:102-103)


class _MatmulOperand:
    """The weight as the GEMM reads it: wq [N,K] 1-byte codes (physical K-major [K,N]), sw [N], optional zp [N] / colsum [N].
    `packed` = weights_dtype when wq is still the stored packed tensor and the GEMM expands it in its prologue."""
    __slots__ = ("wq", "sw", "zp", "colsum", "key", "packed")

    def __init__(self, wq, sw, zp, colsum, key, packed=None):
        self.wq, self.sw, self.zp, self.colsum, self.key, self.packed = wq, sw, zp, colsum, key, packed


def _flat_f32(t):
    return None if t is None else t.reshape(-1).to(torch.float32).contiguous()


_UNTRACKED = 0


def _version(t) -> int:
    """In-place version of a tensor; tensors created under torch.inference_mode() do not track one: every look at such a tensor
    returns a new value, so nothing derived from it is ever taken for current."""
    global _UNTRACKED
    try:
        return t._version
    except RuntimeError:
        _UNTRACKED -= 1
        return _UNTRACKED


class _StoredState:
    """Identity of a layer's stored tensors at the time something was derived from them: a weak reference to each tensor object
    plus its data pointer and in-place version, and the module's generation counter (bumped by SDNQLayer._apply and
    _load_from_state_dict).  data_ptr / _version alone are not enough: CPU-offload re-uploads weights into blocks the caching
    allocator has just recycled (same pointer, version 0) -- those arrive as *new* tensor objects, which the weak references see."""
    __slots__ = ("refs", "marks", "gen")

    def __init__(self, layer, tensors):
        self.refs = tuple(None if t is None else weakref.ref(t) for t in tensors)
        self.marks = tuple(None if t is None else (t.data_ptr(), _version(t)) for t in tensors)
        self.gen = layer.__dict__.get("_sdnq_generation", 0)

    def matches(self, layer, tensors) -> bool:
        if self.gen != layer.__dict__.get("_sdnq_generation", 0) or len(tensors) != len(self.refs):
            return False
        for t, r, m in zip(tensors, self.refs, self.marks):
            if t is None:
                if r is not None:
                    return False
            elif r is None or r() is not t or m != (t.data_ptr(), _version(t)):
                return False
        return True


def _prologue_unpack_ok(d, mm: str, K: int) -> bool:
    """Can K1 expand this layer's stored packed weight in its prologue (sdnq_b200_scaled_mm_packed)?  SDNQ_B200_CACHE_UNPACKED=1
    keeps the round-1 behaviour (expand once with the unpack kernel, cache the N*K-byte operand)."""
    if not d.is_packed or os.environ.get("SDNQ_B200_CACHE_UNPACKED", "0").lower() not in ("0", "false", "no", ""):
        return False
    bits = d.num_bits
    if not 2 <= bits <= 7 or (K * bits) % 128 != 0:              # 16-byte row pitch of the packed rows (TMA)
        return False
    if d.is_integer:
        return mm == "int8"
    info = dtype_dict[d.weights_dtype]
    if mm != "float8_e4m3fn" or info.get("exponent", 9) > 4 or info.get("mantissa", 9) > 3:
        return False
    return True


@torch.no_grad()
def matmul_operand(layer) -> _MatmulOperand:
    """Build (once) and cache the matmul operand of a W8A8 layer.

    reference: per call re_quantize_matmul (dequantizer.py:204-239) / unpack + transpose (linear_int8.py:38-44) / uint8 -> int8
    re-centring (linear_int8.py:45-50).  Weights are frozen, so the result is cached per module and invalidated when the
    stored tensors are replaced (apply_sdnq_options_to_model swaps .data, load_state_dict(assign=True) swaps the Parameter)."""
    d = layer.sdnq_dequantizer
    w, s, z = layer.weight, layer.scale, layer.zero_point
    cached = layer.__dict__.get("_sdnq_mm_cache")
    if cached is not None and cached.key[1] == d.quantized_matmul_dtype and cached.key[0].matches(layer, (w, s, z)):
        return cached
    key = (_StoredState(layer, (w, s, z)), d.quantized_matmul_dtype)
    N, K = d.matmul_nk()
    mm = d.quantized_matmul_dtype
    uint8_mm = d.is_integer_matmul and d.is_unsigned_matmul
    zp = colsum = None
    if d.is_conv and d.is_packed and not d.re_quantize_for_matmul:
        # the reference's conv matmul prologue (conv_int8.py:38-44) transposes the unpacked 4-D weight and fails on this combination too
        raise NotImplementedError("sdnq_b200: packed convolution weights without re-quantisation have no W8A8 kernel")
    if d.re_quantize_for_matmul:
        wq, sw, zp, colsum = d.re_quantize_matmul_raw(w, s, z, want_colsum=uint8_mm)
    elif _prologue_unpack_ok(d, mm, K):
        # row-wise packed weights (2..7-bit integers under an int8 matmul, eXmY minifloats that are subsets of e4m3 under an fp8 one):
        # keep the packed bytes, the GEMM expands them tile by tile in its prologue (no N*K-byte copy)
        op = _MatmulOperand(w.contiguous(), _flat_f32(s), _flat_f32(z), None, key, packed=d.weights_dtype)
        layer.__dict__["_sdnq_mm_cache"] = op
        return op
    elif d.is_packed:
        if d.is_integer:
            wq = ops.unpack(w, d.weights_dtype, (N, K), dtype=torch.int8)          # unsigned codes 0..2^b-1 fit int8 as-is
            zp = _flat_f32(z)
        else:
            wq = ops.unpack(w, d.weights_dtype, (N, K), dtype=torch.float8_e4m3fn)
        sw = _flat_f32(s)
    else:
        if w.dtype not in (torch.int8, torch.uint8, torch.float8_e4m3fn, torch.float8_e5m2):
            raise NotImplementedError(f"sdnq_b200: a {w.dtype} weight cannot be the operand of a {mm} tensor-core matmul")
        wq = ops.physical_nk(w)       # float8_e5m2 codes stay as stored: the GEMM reads them as e5m2 (mixed e4m3 x e5m2 operands)
        sw = _flat_f32(s)
        zp = _flat_f32(z)
        if wq.dtype == torch.uint8:                                                 # uint8 codes -> int8 around 128
            wq = wq.bitwise_xor(128).view(torch.int8)
            zp = torch.add(zp, sw, alpha=128) if zp is not None else torch.mul(sw, 128)
    if uint8_mm and colsum is None:
        colsum = wq.to(torch.int32).sum(dim=1, dtype=torch.int32)
    op = _MatmulOperand(wq.contiguous(), sw, zp, colsum, key)
    layer.__dict__["_sdnq_mm_cache"] = op
    return op


# ---- dequant path: K3 on a side stream -------------------------------------------------------------------------------------
# The dequantised weight depends on nothing but the (frozen) stored tensors, so K3 of layer i can run while the GEMM of layer i-1
# is still executing: it is launched on a per-device side stream and the GEMM waits for it with an event.  At SD-XL bs=1 sizes
# both kernels are latency-bound (~7 us each), so this hides most of K3.  Works eagerly and under CUDA-graph capture (the side
# stream is forked into the capture once and joined at every GEMM).  Safety: if a layer's stored tensors changed identity or
# version since its last call (weights just loaded / converted on the main stream), the side stream first waits for the main
# stream.  SDNQ_B200_DEQUANT_STREAM=0 runs K3 on the caller's stream.
_SIDE_STREAMS: dict = {}


def _side_stream(device):
    st = _SIDE_STREAMS.get(device.index)
    if st is None:
        st = _SIDE_STREAMS[device.index] = torch.cuda.Stream(device=device)
    return st


def _stored_tensors(layer):
    return (layer.weight, layer.scale, layer.zero_point, layer.svd_up, layer.svd_down)


def _dequant_weight(layer, dtype, skip_quantized_matmul):
    d = layer.sdnq_dequantizer
    return d(layer.weight, layer.scale, zero_point=layer.zero_point, svd_up=layer.svd_up, svd_down=layer.svd_down,
             skip_quantized_matmul=skip_quantized_matmul, dtype=dtype)


def _dequant_weight_overlapped(layer, input, skip_quantized_matmul):
    """The dequantised weight of `layer`, produced on the side stream and made visible to the caller's stream."""
    dtype = input.dtype if input.dtype in (torch.float16, torch.bfloat16, torch.float32) else None
    if not input.is_cuda or os.environ.get("SDNQ_B200_DEQUANT_STREAM", "1") in ("0", "false", "no"):
        return _dequant_weight(layer, dtype, skip_quantized_matmul)
    main = torch.cuda.current_stream(input.device)
    side = _side_stream(input.device)
    if dtype == torch.bfloat16 and prefetch.enabled() and prefetch.eligible(layer, dtype):
        # the weights of this layer and of the ones that follow it in the learned call order, dequantised by one launch (prefetch.py)
        W = prefetch.prefetcher(input.device).get(layer, bool(skip_quantized_matmul), main, side)
        if W is not None:
            return W
    # The side stream may run ahead of the main stream only while the stored tensors are provably the ones it has already been
    # ordered after: same tensor objects (weak references), same data pointers and versions, same module generation.  Anything
    # else -- a fresh load, .to(device) of an offloaded module (new Parameter objects in recycled memory), an in-place edit --
    # makes the side stream wait for the main stream first.
    tensors = _stored_tensors(layer)
    seen = layer.__dict__.get("_sdnq_dq_key")
    fresh = seen is None or not seen.matches(layer, tensors)
    if fresh:
        layer.__dict__["_sdnq_dq_key"] = _StoredState(layer, tensors)
    capturing = torch.cuda.is_current_stream_capturing()
    with torch.cuda.stream(side):
        side_capturing = torch.cuda.is_current_stream_capturing()
    if fresh or (capturing and not side_capturing):
        side.wait_stream(main)              # weights may have just been written on the main stream / fork the side stream into the capture
    with torch.cuda.stream(side):
        W = _dequant_weight(layer, dtype, skip_quantized_matmul)
    main.wait_stream(side)
    W.record_stream(main)                   # allocated on the side stream, consumed (and released) on the main one
    return W


def _dequant_linear(layer, input, skip_quantized_matmul):
    return torch.nn.functional.linear(input, _dequant_weight_overlapped(layer, input, skip_quantized_matmul), layer.bias)


def _small_m_packed_ok(self, input) -> bool:
    """K5p applies to rows < 32 of a Linear whose weight is stored packed and / or with group-wise scales (no codebook,
    no tensor-wise scale, 2..8 bits, scale groups that are multiples of 8 columns) and 16-bit activations: the stored bytes are read
    once instead of dequantise + GEMM.  Default; SDNQ_B200_SMALL_M_PACKED=0 restores the reference-shaped dequantise + GEMM."""
    if os.environ.get("SDNQ_B200_SMALL_M_PACKED", "1") in ("0", "false", "no"):
        return False
    d = self.sdnq_dequantizer
    if (d.use_codebook or d.group_size == -2 or self.weight.ndim > 3 or d.is_conv
            or input.dtype not in (torch.bfloat16, torch.float16) or input.numel() == 0):
        return False
    info = dtype_dict[d.weights_dtype]
    N, K = d.matmul_nk()
    if not 2 <= info["num_bits"] <= 8 or K % 16 != 0 or input.shape[-1] != K:
        return False
    per_row = self.scale.numel() // N
    return per_row >= 1 and self.scale.numel() == N * per_row and K % per_row == 0 and (K // per_row) % 8 == 0


def _small_m_packed_linear(self, x):
    d = self.sdnq_dequantizer
    N, K = d.matmul_nk()
    if d.use_hadamard:      # x @ (W_rot H)^T = (x H) @ W_rot^T: rotate the (tiny) activation with K2 instead of un-rotating the weight
        x = ops.act_quant(x, "int8", hadamard_group=d.hadamard_group_size, want_x_rot=True)[4].view(x.shape)
    bias = self.bias
    if self.svd_up is not None:
        # W = dequant + svd_up @ svd_down (dequantizer.py:69-79)  =>  y = x @ dequant^T + (x @ svd_down^T) @ svd_up^T: the rank-r term
        # of these few rows is two skinny library GEMMs handed to the kernel as an [M,N] bias.  Factors are stored [N,r] / [r,K], or
        # transposed ([r,N] / [K,r]) when the layer is in matmul layout (quantizer.py:164-167).
        up, down = self.svd_up, self.svd_down
        if not d.use_quantized_matmul:
            up, down = up.t(), down.t()                     # -> [r,N], [K,r]
        x2 = x.reshape(-1, K).to(down.dtype)
        low = torch.mm(x2, down)
        bias = torch.mm(low, up) if bias is None else torch.addmm(bias.to(down.dtype), low, up)
    return ops.linear_small_m_packed(x, self.weight, d.weights_dtype, self.scale, self.zero_point, N, K, bias=bias)


# ---- K6: the dequant path in one launch ("W4A16") ----------------------------------------------------------------------------
def w4a16_enabled() -> bool:
    """SDNQ_B200_W4A16=1 routes 4-bit dequant-path layers through the one-launch W4A16 kernel (K6) instead of dequantise (K3 on
    the side stream) + library GEMM.  Off by default: measured on B200 (profiles/r02_w4a16.md) the fused kernel re-dequantises
    the weight tile for every 128-row strip of activations and is bound by the L2 -> SM operand traffic of its 128 x 128 tiles;
    at the SD-XL / FLUX shapes of BASELINE.json (M >= 77 rows) dequantise-once + GEMM is 1.4 - 2x faster."""
    return os.environ.get("SDNQ_B200_W4A16", "0") not in ("0", "false", "no", "")


def _w4a16_ok(self, input) -> bool:
    """K6 applies to 32 or more rows of a dequant-path Linear stored as packed int4 / uint4 with row-wise scales or scale groups
    that are multiples of 32 columns (no codebook, no tensor-wise scale), 16-bit activations, K % 64 == 0, N % 8 == 0 and, if
    the layer has SVD factors, rank 16 / 32 / 64 in the activation dtype."""
    d = self.sdnq_dequantizer
    if (not w4a16_enabled() or d.is_conv or d.use_codebook or d.group_size == -2 or not d.is_integer or d.num_bits != 4
            or input.dtype not in (torch.bfloat16, torch.float16) or self.weight.ndim > 3 or not input.is_cuda):
        return False
    N, K = d.matmul_nk()
    if K % 64 != 0 or N % 8 != 0 or input.shape[-1] != K:
        return False
    per_row = self.scale.numel() // N
    if per_row < 1 or self.scale.numel() != N * per_row or K % per_row != 0 or (per_row > 1 and (K // per_row) % 32 != 0):
        return False
    if self.svd_up is not None:
        r = min(self.svd_up.shape)
        if r not in (16, 32, 64) or self.svd_up.dtype != input.dtype or self.svd_down.dtype != input.dtype:
            return False
    return True


def _svd_operands(self):
    """The SVD factors as K6 reads them -- svd_down [r,K] and svd_up [N,r], both row-major -- built once per layer from the stored
    tensors (matmul-off layout: svd_up [N,r], svd_down [r,K] kept K-major, i.e. physically [K,r]; matmul-on layout: both
    transposed, quantizer.py:164-167) and rebuilt when those are replaced."""
    up, down = self.svd_up, self.svd_down
    if up is None:
        return None, None
    cached = self.__dict__.get("_sdnq_svd_cache")
    if cached is not None and cached[0].matches(self, (up, down)):
        return cached[1], cached[2]
    if self.sdnq_dequantizer.use_quantized_matmul:
        up, down = up.t(), down.t()                              # -> [N,r], [r,K]
    down_rk, up_nr = down.contiguous(), up.contiguous()
    self.__dict__["_sdnq_svd_cache"] = (_StoredState(self, (self.svd_up, self.svd_down)), down_rk, up_nr)
    return down_rk, up_nr


def _svd_mm_operands(self, x_dtype):
    """The SVD factors as K7 / K1 read them on the W8A8 path: svd_down [r',K] and svd_up [N,r'] row-major in the activation dtype, the
    rank padded with zero rows / columns to r' = 16, 32 or 64 (the tcgen05 kind::f16 k-step and the TMA swizzle spans).  None when the
    kernels do not apply (rank > 64, factors not stored in the 16-bit activation dtype): the caller then builds the reference's
    dense [M,N] bias instead."""
    up, down = self.svd_up, self.svd_down
    if (x_dtype not in (torch.bfloat16, torch.float16) or up.dtype != x_dtype or down.dtype != x_dtype or min(up.shape) > 64
            or max(down.shape) % 16 != 0 or os.environ.get("SDNQ_B200_SVD_FUSED", "1") in ("0", "false", "no")):
        return None
    cached = self.__dict__.get("_sdnq_svd_mm_cache")
    if cached is not None and cached[0].matches(self, (up, down)):
        return cached[1], cached[2]
    down_rk, up_nr = _svd_operands(self)
    r = down_rk.shape[0]
    rp = 16 if r <= 16 else 32 if r <= 32 else 64
    if rp != r:
        down_rk = torch.nn.functional.pad(down_rk, (0, 0, 0, rp - r)).contiguous()
        up_nr = torch.nn.functional.pad(up_nr, (0, rp - r)).contiguous()
    self.__dict__["_sdnq_svd_mm_cache"] = (_StoredState(self, (up, down)), down_rk, up_nr)
    return down_rk, up_nr


def _svd_bias2d(self, x_rot):
    """The reference's SVD term as a dense bias (linear_int8.py:57-62): used only where K7 + the rank-r accumulate do not apply."""
    down, up = self.svd_down, self.svd_up
    low = torch.mm(x_rot.to(down.dtype), down)
    return torch.mm(low, up) if self.bias is None else torch.addmm(self.bias.to(down.dtype), low, up)


def _w4a16_linear(self, x):
    d = self.sdnq_dequantizer
    N, K = d.matmul_nk()
    if d.use_hadamard:      # x @ ((Wq + up down) H)^T = (x H) @ (Wq + up down)^T: rotate the activations with K2, not the weight
        x = ops.act_quant(x, "int8", hadamard_group=d.hadamard_group_size, want_x_rot=True)[4].view(x.shape)
    down_rk, up_nr = _svd_operands(self)
    return ops.linear_w4a16(x, self.weight, d.weights_dtype, self.scale, self.zero_point, N, K, bias=self.bias,
                            svd_down_rk=down_rk, svd_up_nr=up_nr)


@torch.no_grad()
def quantized_linear_forward(self, input: torch.Tensor) -> torch.Tensor:
    rows = input.numel() // max(input.shape[-1], 1)
    if rows < SMALL_M:
        if _small_m_packed_ok(self, input):
            return _small_m_packed_linear(self, input)
    elif _w4a16_ok(self, input):
        return _w4a16_linear(self, input)
    else:
        group = self.__dict__.get("_sdnq_siblings")
        if type(group) is siblings.DequantSiblingGroup and siblings.siblings_enabled():      # to_q / to_k / to_v ...: one batched GEMM over their dequantised weights
            out = group.forward(self, input)
            if out is not None:
                return out.view(*input.shape[:-1], out.shape[-1])
    return _dequant_linear(self, input, skip_quantized_matmul=False)


def _small_m_gemv_ok(self, input) -> bool:
    """K5 applies when the stored weight is itself the row-wise 8-bit matmul operand (int8 / uint8 / float8_e4m3fn, no re-quantise,
    no packing, no SVD) and the activations are 16-bit.  SDNQ_B200_SMALL_M_GEMV=0 keeps the reference's dequantise + GEMM."""
    d = self.sdnq_dequantizer
    return (not d.re_quantize_for_matmul and not d.is_packed and self.svd_up is None and input.dtype in (torch.bfloat16, torch.float16)
            and self.weight.dtype in (torch.int8, torch.uint8, torch.float8_e4m3fn, torch.float8_e5m2)
            and input.shape[-1] % 16 == 0 and input.numel() > 0 and os.environ.get("SDNQ_B200_SMALL_M_GEMV", "1") not in ("0", "false", "no"))


# ---- K2 reuse: sibling projections quantise the same activations once ------------------------------------------------------
# to_q / to_k / to_v of an attention block (FLUX single blocks: proj_mlp as well) and the cross-attention to_k / to_v of *every*
# block (all fed the same encoder_hidden_states) call their forward with the very same input tensor, and the reference rotates
# and row-quantises it again for each of them (linear_int8.py:55-63).  Here the result of K2 is kept -- per device and stream,
# the newest entry plus a few small older ones -- and reused when a W8A8 forward on that stream presents the same tensor: same storage pointer, shape,
# strides, dtype and in-place version, with a strong reference to the cached input held so that its memory cannot be recycled
# under a new tensor.  The key carries the stream's CUDA-graph capture id: nothing computed eagerly is baked into a graph and
# nothing from one capture is used in another.  SDNQ_B200_ACT_CACHE=0 turns it off.
_ACT_CACHE: dict = {}
_ACT_CACHE_ENTRIES = 4                   # per (device, stream): the newest entry plus a few small ones
_ACT_CACHE_SMALL_BYTES = 8 << 20         # older entries are kept only while they pin less than this in total (e.g. the text
                                         # encoder states every cross-attention block projects: 77 x 2048 for SD-XL)


def _act_cache_on() -> bool:
    return os.environ.get("SDNQ_B200_ACT_CACHE", "1") not in ("0", "false", "no")


def quantized_activations(x2: torch.Tensor, mm: str, hg: int, want_rowsum: bool, want_x_rot: bool):
    """K2 on x2 [M,K] -> (xq, sx, zx, rowsum, x_rot), shared between forwards that present the same tensor."""
    if not _act_cache_on() or not x2.is_cuda:       # (a CPU tensor raises inside ops: there is no CPU path)
        return ops.act_quant(x2, mm, hadamard_group=hg, want_rowsum=want_rowsum, want_x_rot=want_x_rot)
    stream = torch.cuda.current_stream(x2.device).cuda_stream
    slot = (x2.device.index, stream)
    key = (x2.data_ptr(), _version(x2), tuple(x2.shape), tuple(x2.stride()), x2.dtype, mm, int(hg), ops.capture_id(stream))
    entries = _ACT_CACHE.setdefault(slot, [])
    for i, (k, _, out) in enumerate(entries):
        if k == key and (not want_rowsum or out[3] is not None) and (not want_x_rot or out[4] is not None):
            if i:
                entries.insert(0, entries.pop(i))
            return out
    out = ops.act_quant(x2, mm, hadamard_group=hg, want_rowsum=want_rowsum, want_x_rot=want_x_rot)
    # the newest entry always stays; older ones only while they are small (x2 is kept alive while its entry is, so that its memory
    # cannot be recycled under a new tensor) and belong to the same graph capture (or to none)
    kept, held = [(key, x2, out)], 0
    for e in entries:
        nbytes = 2 * e[1].numel() * e[1].element_size()
        if len(kept) < _ACT_CACHE_ENTRIES and e[0][-1] == key[-1] and held + nbytes <= _ACT_CACHE_SMALL_BYTES:
            kept.append(e)
            held += nbytes
    entries[:] = kept
    return out


def _rows(input: torch.Tensor) -> torch.Tensor:
    return input.reshape(-1, input.shape[-1])


def _w8a8_forward(self, input: torch.Tensor) -> torch.Tensor:
    d = self.sdnq_dequantizer
    if input.numel() // input.shape[-1] < SMALL_M:
        if _small_m_gemv_ok(self, input):
            # rows < 32 (linear_int8.py:102-103): the reference dequantises the whole weight and runs F.linear; K5 reads the codes
            # once instead.  Rotated layers: x @ (W_rot H)^T = (x H) @ W_rot^T, so rotate the (tiny) activation, not the weight.
            op = matmul_operand(self)
            x = input
            if d.use_hadamard:
                x = ops.act_quant(input, d.quantized_matmul_dtype, hadamard_group=d.hadamard_group_size, want_x_rot=True)[4].view(input.shape)
            return ops.linear_small_m(x, op.wq, op.sw, zp=op.zp, bias=self.bias)
        if (d.re_quantize_for_matmul or d.is_packed) and _small_m_packed_ok(self, input):
            # packed / group-wise weights: K5p reads the stored bytes once
            return _small_m_packed_linear(self, input)
        return _dequant_linear(self, input, skip_quantized_matmul=True)
    op = matmul_operand(self)
    mm = d.quantized_matmul_dtype
    hg = d.hadamard_group_size if d.use_hadamard else 0
    x2 = _rows(input)
    if self.svd_up is None:
        group = self.__dict__.get("_sdnq_siblings")
        if type(group) is siblings.SiblingGroup and siblings.siblings_enabled():      # to_q / to_k / to_v ...: one K2 + one grouped K1 launch for all of them
            out = group.forward(self, x2, input.dtype)
            if out is not None:
                return out.view(*input.shape[:-1], out.shape[-1])
        if op.packed is None and not _act_cache_on():       # one C call: K2 into the per-stream workspace + K1
            return ops.linear_w8a8(input, op.wq, mm, op.sw, bias=self.bias, zp=op.zp, colsum=op.colsum, hadamard_group=hg, out_dtype=input.dtype)
        xq, sx, zx, rowsum, _ = quantized_activations(x2, mm, hg, op.zp is not None, False)
        if op.packed is not None:
            out = ops.scaled_mm_packed(xq, op.wq, op.packed, d.original_shape[0], sx, op.sw, self.bias, input.dtype, rowsum=rowsum, zp=op.zp)
        else:
            out = ops.scaled_mm(xq, op.wq, sx, op.sw, self.bias, input.dtype, rowsum=rowsum, zp=op.zp, colsum=op.colsum, zx=zx)
        return out.view(*input.shape[:-1], out.shape[-1])
    # SVD branch (linear_int8.py:57-62): bias2d = bias + (x_rot @ svd_down[K,r]) @ svd_up[r,N] on the rotated, un-quantised activations
    # in the SVD dtype.  K7 computes low = x_rot @ svd_down, K1 accumulates low @ svd_up per output tile on the tensor cores.
    # (the rank-r accumulate exists for unpacked and 4-bit packed operands; other packed widths take the dense-bias form)
    svd = _svd_mm_operands(self, input.dtype) if (op.packed is None or (d.is_integer and d.num_bits == 4)) else None
    need_rot = hg != 0 or svd is None                       # without rotation x_rot is x itself: K2 does not write a copy of it
    xq, sx, zx, rowsum, x_rot = quantized_activations(x2, mm, hg, op.zp is not None, need_rot)
    if svd is None:
        bias2d = _svd_bias2d(self, x_rot)
        if op.packed is not None:
            out = ops.scaled_mm_packed(xq, op.wq, op.packed, d.original_shape[0], sx, op.sw, bias2d, input.dtype, rowsum=rowsum, zp=op.zp)
        else:
            out = ops.scaled_mm(xq, op.wq, sx, op.sw, bias2d, input.dtype, rowsum=rowsum, zp=op.zp, colsum=op.colsum, zx=zx)
        return out.view(*input.shape[:-1], out.shape[-1])
    down_rk, up_nr = svd
    low = ops.svd_low(x_rot if hg != 0 else x2, down_rk)
    out = ops.scaled_mm_svd(xq, op.wq, sx, op.sw, low, up_nr, self.bias, input.dtype, rowsum=rowsum, zp=op.zp, colsum=op.colsum, zx=zx,
                            packed_dtype=op.packed, N=d.original_shape[0])
    return out.view(*input.shape[:-1], out.shape[-1])


@torch.no_grad()
def quantized_linear_forward_int8_matmul(self, input: torch.Tensor) -> torch.Tensor:
    return _w8a8_forward(self, input)


@torch.no_grad()
def quantized_linear_forward_uint8_matmul(self, input: torch.Tensor) -> torch.Tensor:
    return _w8a8_forward(self, input)


@torch.no_grad()
def quantized_linear_forward_fp8_matmul(self, input: torch.Tensor) -> torch.Tensor:
    return _w8a8_forward(self, input)


def _unsupported(name: str, what: str) -> Callable:
    def forward(self, *args, **kwargs):
        raise NotImplementedError(f"sdnq_b200: {what} has no sm_100a kernel yet; refusing to fall back to an eager path")
    forward.__name__ = name
    return forward


quantized_linear_forward_fp16_matmul = _unsupported("quantized_linear_forward_fp16_matmul", "the float16 quantized matmul (quantized_matmul_dtype='float16')")
quantized_conv_forward_fp16_matmul = _unsupported("quantized_conv_forward_fp16_matmul", "the float16 quantized conv matmul (quantized_matmul_dtype='float16')")


@torch.no_grad()
def quantized_embedding_forward(self, input: torch.Tensor) -> torch.Tensor:
    """Quantized embedding (quant_embedding=True; reference layers/embedding/forward.py:14-104): the reference unpacks the whole table,
    indexes weight / scale / zero point / svd_up with the token ids and dequantises those rows; K3 reads the stored rows of the
    ids directly (the table is never unpacked) and applies scalar_embed_scale on the way out."""
    d = self.sdnq_dequantizer
    shape = tuple(d.result_shape) if d.result_shape is not None else tuple(d.original_shape)
    if len(shape) != 2:
        raise NotImplementedError(f"sdnq_b200: embedding tables are 2-D (got shape {shape})")
    V, D = shape
    embed_scale = getattr(self, "scalar_embed_scale", None)
    if torch.is_tensor(embed_scale):
        embed_scale = float(embed_scale)
    un_rotate = d.hadamard_group_size if d.use_hadamard else 0
    return ops.embedding(self.weight, d.weights_dtype, self.scale, self.zero_point, V, D, d.group_size, input, d.result_dtype,
                         svd_up=self.svd_up, svd_down=self.svd_down, hadamard_group=un_rotate, use_codebook=d.use_codebook,
                         embed_scale=1.0 if embed_scale is None else embed_scale)


# ------------------------------------------------------------------------------------------------ convolutions
def _conv_dense_weight(self, skip_quantized_matmul, input=None):
    if input is not None and input.dtype == self.sdnq_dequantizer.result_dtype:
        return _dequant_weight_overlapped(self, input, skip_quantized_matmul)      # K3c on the side stream, like the Linear dequant path
    d = self.sdnq_dequantizer
    return d(self.weight, self.scale, zero_point=self.zero_point, svd_up=self.svd_up, svd_down=self.svd_down,
             skip_quantized_matmul=skip_quantized_matmul)


@torch.no_grad()
def quantized_conv_forward(self, input: torch.Tensor) -> torch.Tensor:
    """K3 (broadcast dequant of the conv weight) -> the library convolution (reference layers/conv/forward.py:79-81)."""
    return self._conv_forward(input, _conv_dense_weight(self, False, input), self.bias)


def _conv_transpose_forward(nd, fn):
    @torch.no_grad()
    def forward(self, input: torch.Tensor, output_size=None) -> torch.Tensor:
        output_padding = self._output_padding(input, output_size, self.stride, self.padding, self.kernel_size, nd, self.dilation)
        return fn(input, _conv_dense_weight(self, False, input), self.bias, self.stride, self.padding, output_padding, self.groups, self.dilation)
    forward.__name__ = f"quantized_conv_transpose_{nd}d_forward"
    return forward


# reference layers/conv/forward.py:84-99
quantized_conv_transpose_1d_forward = _conv_transpose_forward(1, torch.nn.functional.conv_transpose1d)
quantized_conv_transpose_2d_forward = _conv_transpose_forward(2, torch.nn.functional.conv_transpose2d)
quantized_conv_transpose_3d_forward = _conv_transpose_forward(3, torch.nn.functional.conv_transpose3d)


def _pair(v, n):
    return (int(v),) * n if isinstance(v, int) else tuple(int(i) for i in v)


def conv_matmul_unsupported(module) -> str | None:
    """Why the quantized conv matmul of this convolution has no sm_100a kernel (None = it has one)."""
    if getattr(module, "groups", 1) != 1:
        return "a grouped convolution"
    if len(tuple(getattr(module, "kernel_size", (1, 1)))) > 2:
        return "Conv3d"
    if isinstance(getattr(module, "padding", 0), str):
        return "string padding ('same' / 'valid')"
    return None


def _w8a8_conv_forward(self, input: torch.Tensor) -> torch.Tensor:
    """conv_{int8,uint8,fp8}_matmul (reference layers/conv/conv_int8.py:17-125, conv_uint8.py, conv_fp8.py): the convolution as
    one W8A8 GEMM over the im2col view.  Here: K2 gathers each im2col row straight from the input (the [M, C*kh*kw] bf16
    matrix of F.unfold is never written), rotates / row-quantises it; K1 contracts it with the row-wise weight; the
    [B, H_out, W_out, N] result is returned in the reference's NCHW-contiguous form."""
    d = self.sdnq_dequantizer
    if input.numel() / input.shape[2] < SMALL_M:                                  # conv_int8.py:95-96
        return self._conv_forward(input, _conv_dense_weight(self, True, input), self.bias)
    why = conv_matmul_unsupported(self)
    if why is not None:
        # sdnq_quantize_layer keeps such layers on the dequant path when it quantises them (with a warning); this is reached only by
        # a checkpoint that was written with the quantized conv matmul on for them
        raise NotImplementedError(f"sdnq_b200: {why} has no W8A8 conv kernel; call apply_sdnq_options_to_model(model, "
                                  "use_quantized_matmul=False) or list the layer in modules_to_not_use_matmul")
    nd = input.ndim - 2
    ksz, stride, padding, dilation = (_pair(v, nd) for v in (self.kernel_size, self.stride, self.padding, self.dilation))
    x4 = input
    if self.padding_mode != "zeros":                                              # process_conv_input (layers/conv/forward.py:53-55)
        x4 = torch.nn.functional.pad(input, self._reversed_padding_repeated_twice, mode=self.padding_mode)
        padding = (0,) * nd
    if nd == 1:                                                                   # get_conv_args: conv1d = conv2d with H = 1
        x4 = x4.unsqueeze(2)
        ksz, stride, padding, dilation = (1, ksz[0]), (1, stride[0]), (0, padding[0]), (1, dilation[0])
    op = matmul_operand(self)
    mm = d.quantized_matmul_dtype
    hg = d.hadamard_group_size if d.use_hadamard else 0
    svd = self.svd_up is not None
    xq, sx, zx, rowsum, x_rot, (B, Ho, Wo) = ops.conv_act_quant(x4, ksz, stride, padding, dilation, mm, hadamard_group=hg,
                                                                want_rowsum=op.zp is not None, want_x_rot=svd)
    svd_ops = _svd_mm_operands(self, input.dtype) if svd and (op.packed is None or (d.is_integer and d.num_bits == 4)) else None
    if svd_ops is not None:                                                       # conv_int8.py:56-61 as K7 + the rank-r accumulate in K1
        low = ops.svd_low(x_rot, svd_ops[0])
        out = ops.scaled_mm_svd(xq, op.wq, sx, op.sw, low, svd_ops[1], self.bias, input.dtype, rowsum=rowsum, zp=op.zp, colsum=op.colsum, zx=zx,
                                packed_dtype=op.packed, N=d.original_shape[0])
    else:
        bias = _svd_bias2d(self, x_rot) if svd else self.bias
        if op.packed is not None:
            out = ops.scaled_mm_packed(xq, op.wq, op.packed, d.original_shape[0], sx, op.sw, bias, input.dtype, rowsum=rowsum, zp=op.zp)
        else:
            out = ops.scaled_mm(xq, op.wq, sx, op.sw, bias, input.dtype, rowsum=rowsum, zp=op.zp, colsum=op.colsum, zx=zx)
    N = out.shape[-1]
    y = ops.rows_to_nchw(out, B, Ho * Wo)                                         # [B*L, N] -> [B, N, L] (conv_int8.py:83-89)
    return y.view(B, N, Wo) if nd == 1 else y.view(B, N, Ho, Wo)


@torch.no_grad()
def quantized_conv_forward_int8_matmul(self, input: torch.Tensor) -> torch.Tensor:
    return _w8a8_conv_forward(self, input)


@torch.no_grad()
def quantized_conv_forward_uint8_matmul(self, input: torch.Tensor) -> torch.Tensor:
    return _w8a8_conv_forward(self, input)


@torch.no_grad()
def quantized_conv_forward_fp8_matmul(self, input: torch.Tensor) -> torch.Tensor:
    return _w8a8_conv_forward(self, input)


def get_forward_func(layer_class_name: str, quantized_matmul_dtype: str, use_quantized_matmul: bool) -> Callable:
    """reference forward.py:6-57."""
    if layer_class_name in embedding_types:
        return quantized_embedding_forward
    if layer_class_name in conv_types:
        if not use_quantized_matmul:
            return quantized_conv_forward
        mm = dtype_dict[quantized_matmul_dtype]
        if mm["is_integer"]:
            return quantized_conv_forward_uint8_matmul if mm["is_unsigned"] else quantized_conv_forward_int8_matmul
        return quantized_conv_forward_fp8_matmul if mm["num_bits"] == 8 else quantized_conv_forward_fp16_matmul
    if layer_class_name in conv_transpose_types:
        return {"1": quantized_conv_transpose_1d_forward, "2": quantized_conv_transpose_2d_forward,
                "3": quantized_conv_transpose_3d_forward}[layer_class_name[-2]]
    if not use_quantized_matmul:
        return quantized_linear_forward
    mm = dtype_dict[quantized_matmul_dtype]
    if mm["is_integer"]:
        return quantized_linear_forward_uint8_matmul if mm["is_unsigned"] else quantized_linear_forward_int8_matmul
    return quantized_linear_forward_fp8_matmul if mm["num_bits"] == 8 else quantized_linear_forward_fp16_matmul
