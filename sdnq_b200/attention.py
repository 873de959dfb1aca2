"""SDNQ quantized attention on the sm_100a kernels: the host-side mirror of the reference's
`kernels/triton_atten.py` -- `quantize_attn` (:443-487), `get_attn_inputs` (:490-536) and the SDPA drop-in `sdnq_triton_atten`
(:540-615) -- over K2 (`act_quant`: Hadamard rotation + per-row int8 / fp8 quantisation of q and k), the smooth-K kernel and K9
(`attention_fwd`: tcgen05 QK^T on the 1-byte codes, online softmax, 16-bit or 1-byte P.V).

Covered: int8 / float8_e4m3fn Q.K^T (`matmul_dtype`), unquantised P.V (`pv_matmul_dtype=None`, the reference's default) and
int8 / float8_e4m3fn P.V (`pv_matmul_dtype`: v quantised per key over the head dim, P per row and key tile inside the kernel, the
rotated output un-rotated afterwards), smooth-K, Hadamard rotation of q / k (/ v), boolean / additive masks, causal, grouped-query
heads, the log-sum-exp output.  Anything else the reference accepts (fp16 codes, fp16 accumulation, unquantised Q.K^T) raises
NotImplementedError: there is no fallback that would silently run another implementation."""
import torch
import torch.nn.functional as F

from . import ops

_DISABLED = {None, "none", "no", "disabled"}


def _next_pow2(n: int) -> int:
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


def get_hadamard_group_size(channel_size: int, group_size: int):
    """quant_utils.py:212-218"""
    group_size = _next_pow2(min(channel_size, group_size))
    while channel_size % group_size != 0:
        group_size //= 2
    return group_size >= 4, group_size


def _mm_dtype(matmul_dtype):
    if matmul_dtype in {"auto", "enabled", "uint8", "int8"}:                  # triton_atten.py:452-453
        return "int8"
    if matmul_dtype in {"fp8", "float8_e4m3fn"}:
        return "float8_e4m3fn"
    raise NotImplementedError(f"sdnq_b200 attention: matmul_dtype {matmul_dtype!r} has no sm_100a kernel (int8 and float8_e4m3fn do)")


def _pv_dtype(pv_matmul_dtype):
    """None when P.V stays unquantised (triton_atten.py:478: None / "auto" / "none" / "no" / "disabled"), else the code dtype"""
    if pv_matmul_dtype in _DISABLED or pv_matmul_dtype == "auto":
        return None
    if pv_matmul_dtype in {"enabled", "uint8", "int8"}:                      # :454-455
        return "int8"
    if pv_matmul_dtype in {"fp8", "float8_e4m3fn"}:
        return "float8_e4m3fn"
    raise NotImplementedError(f"sdnq_b200 attention: pv_matmul_dtype {pv_matmul_dtype!r} has no sm_100a kernel (int8 and float8_e4m3fn do)")


def _quantize_v(v, pv, G):
    """triton_atten.py:478-483: v rotated like q / k when they are, then quantised per key over the head dim -> (codes, scale [Z,VH,KN])"""
    if G == 0 and v.shape[-1] in (16, 32, 64, 128, 256):
        return ops.attn_quant(v, pv)
    if G and v.shape[-1] % G != 0:
        raise ValueError(f"sdnq_b200 attention: the Hadamard group {G} of q / k does not divide v's head dim {v.shape[-1]}")
    v_q, v_scale, _, _, _ = ops.act_quant(v, pv, hadamard_group=G)
    return v_q.view(v.shape), v_scale.view(v.shape[:-1])


def quantize_attn(q, k, v, smooth_k: bool = True, hadamard_group_size: int = 0, matmul_dtype: str = "int8", pv_matmul_dtype=None):
    """triton_atten.py:443-487 for the covered cases.  q [Z,H,QN,HD], k [Z,KH,KN,HD], v [Z,VH,KN,HDV] (16-bit or f32).
    hadamard_group_size: 0 = no rotation, else the group the caller resolved (sdnq_triton_atten :560-566).
    -> (q_q, q_scale [Z,H,QN], k_q, k_scale [Z,KH,KN], v | v_q, None | v_scale [Z,VH,KN])"""
    if matmul_dtype in _DISABLED:
        raise NotImplementedError("sdnq_b200 attention: unquantised Q.K^T is not built (use torch SDPA)")
    pv = _pv_dtype(pv_matmul_dtype)
    mm = _mm_dtype(matmul_dtype)
    G = int(hadamard_group_size)
    HD = q.shape[-1]
    v_scale = None
    if pv is not None:
        if G and v.dtype != q.dtype:
            v = v.to(q.dtype)                                                  # :480: the rotation runs in its own (= q's) dtype
        v, v_scale = _quantize_v(v, pv, G)
    if G == 0 and HD == k.shape[-1] and HD in (16, 32, 64, 128, 256):
        # no rotation: one row-quantiser launch per operand (the channel means of smooth-K are subtracted inside it)
        q_q, q_scale = ops.attn_quant(q, mm)
        k_q, k_scale = ops.attn_quant(k, mm, smooth=smooth_k)
        return q_q, q_scale, k_q, k_scale, v, v_scale
    if smooth_k:
        # :456-461 (k - mean in f32); :463-466: with a rotation the result is cast to the rotation's dtype (= q's) first
        k = ops.smooth_k(k, q.dtype if G else torch.float32)
    elif G and k.dtype != q.dtype:
        k = k.to(q.dtype)
    q_q, q_scale, _, _, _ = ops.act_quant(q, mm, hadamard_group=G)
    k_q, k_scale, _, _, _ = ops.act_quant(k, mm, hadamard_group=G)
    return q_q.view(q.shape), q_scale.view(q.shape[:-1]), k_q.view(k.shape), k_scale.view(k.shape[:-1]), v, v_scale


def sdnq_attention(query, key, value, attn_mask=None, dropout_p: float = 0.0, is_causal: bool = False, scale=None, enable_gqa: bool = False,
                   smooth_k: bool = True, use_hadamard: bool = False, hadamard_group_size: int = 256, matmul_dtype: str = "int8",
                   pv_matmul_dtype=None, do_quantize: bool = True, use_fp16_accum: bool = False, out_dtype=None, return_lse: bool = False):
    """`sdnq_triton_atten` (triton_atten.py:540-615): scaled-dot-product attention over [Z, heads, tokens, head_dim] tensors with
    quantised Q.K^T.  Same arguments; `return_lse=True` additionally returns the base-2 log-sum-exp [Z,H,QN] (the reference's
    `return_backward` bundle is training-side and out of scope)."""
    if not do_quantize:
        raise NotImplementedError("sdnq_b200 attention: do_quantize=False (plain flash attention) is not built (use torch SDPA)")
    if use_fp16_accum:
        raise NotImplementedError("sdnq_b200 attention: use_fp16_accum has no tcgen05 equivalent (accumulators are s32 / f32 in TMEM)")
    QHD, KHD, VHD = query.shape[-1], key.shape[-1], value.shape[-1]
    if out_dtype is None:
        out_dtype = query.dtype                                                # :510-511
    sm_scale = QHD ** -0.5 if scale is None else float(scale)                 # :512-513 (before padding)
    G = 0
    if use_hadamard and matmul_dtype not in _DISABLED:                       # :560-566
        channel = _next_pow2(min(QHD, KHD))
        ok, g = get_hadamard_group_size(channel, min(hadamard_group_size, channel))
        G = g if ok else 0
    if QHD != _next_pow2(QHD) or QHD < 16:                                    # :514-519 (the kernel needs a multiple of 16)
        query = F.pad(query, (0, max(_next_pow2(QHD), 16) - QHD))
    if KHD != _next_pow2(KHD) or KHD < 16:
        key = F.pad(key, (0, max(_next_pow2(KHD), 16) - KHD))
    vpad = max(_next_pow2(VHD), 64)
    if vpad != VHD:
        value = F.pad(value, (0, vpad - VHD))
    pv = _pv_dtype(pv_matmul_dtype)
    if pv is None and value.dtype not in (torch.bfloat16, torch.float16):
        value = value.to(query.dtype if query.dtype in (torch.bfloat16, torch.float16) else torch.bfloat16)
    q_q, q_scale, k_q, k_scale, value, v_scale = quantize_attn(query, key, value, smooth_k=smooth_k, hadamard_group_size=G, matmul_dtype=matmul_dtype,
                                                             pv_matmul_dtype=pv_matmul_dtype)
    out, lse = ops.attention_fwd(q_q, k_q, value, q_scale, k_scale, attn_mask=attn_mask, is_causal=is_causal, sm_scale=sm_scale, out_dtype=out_dtype,
                                 return_lse=return_lse, v_scale=v_scale)
    if G and pv is not None:
        # :604-607: the output of P.(v H) is rotated back (H is symmetric and orthonormal: the same rotation); K2's x_rot output
        out = ops.act_quant(out, "int8", hadamard_group=G, want_x_rot=True)[4].view(out.shape)
    out = out[..., :VHD]
    return (out, lse) if return_lse else out


sdnq_triton_atten = sdnq_attention      # the reference's name for the same entry point
