"""CPU oracle for the SDNQ quantized-Linear forward  --  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch numpy restatement of the reference algorithm for the
hot path.  It is the checker, never the product: only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  Nothing under `sdnq_b200/` imports it.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function here
against fixtures produced by running the unmodified reference in the authoring
container (`tests/golden/generate.py`, flags SDNQ_USE_CONTIGUOUS_MM=0
SDNQ_ALLOW_FP8_MM=1 = the flag state the reference resolves to on a B200).
The reference ships no tests / golden vectors of its own.

Conventions
  * every array is numpy; "bf16 tensors" are float32 arrays whose values are
    exactly representable in bfloat16 (`bf16_round`), fp8 tensors are float32
    arrays holding e4m3fn-representable values (`e4m3fn_round`) -- numpy has
    neither type, and holding the value in f32 keeps the arithmetic explicit;
  * arrays are *logical* (a K-major `[K,N]` weight is just `phys.T`);
  * reference citations are relative to /root/reference/src/sdnq/.
"""
from __future__ import annotations

import re

import numpy as np

F32 = np.float32


# =============================================================================
# dtype descriptors  (reference: common.py:16-267 -- regenerated from the naming rule)
# =============================================================================
_FLOAT_RE = re.compile(r"^float(\d+)_e(\d+)m(\d+)(fnu|fn)(_sdnq)?$")
_ALIASES = {
    "fp8": "float8_e4m3fn", "fp16": "float16", "bf16": "bfloat16", "fp32": "float32", "bool": "uint1", "int1": "uint1",
}


def dtype_info(name: str) -> dict:
    """min / max / num_bits / is_unsigned / is_integer / is_packed / exponent / mantissa of a storage dtype."""
    name = _ALIASES.get(name, name)
    m = re.match(r"^(u?)int(\d+)$", name)
    if m:
        unsigned, bits = m.group(1) == "u", int(m.group(2))
        lo, hi = (0, 2 ** bits - 1) if unsigned else (-(2 ** (bits - 1)), 2 ** (bits - 1) - 1)
        if unsigned and 9 <= bits <= 15:
            hi = 2 ** bits  # upstream table quirk (common.py:40-46 lists max = 2^bits for uint9..uint15); kept for parity
        return dict(name=name, min=lo, max=hi, num_bits=bits, is_unsigned=unsigned, is_integer=True,
                    is_packed=bits not in (8, 16, 32), exponent=0, mantissa=bits - (0 if unsigned else 1))
    native = {
        "float32": (8, 23, 3.40282e38), "bfloat16": (8, 7, 3.38953e38), "float16": (5, 10, 65504.0),
        "float8_e4m3fn": (4, 3, 448.0), "float8_e5m2": (5, 2, 57344.0),
    }
    if name in native:
        e, mm, mx = native[name]
        return dict(name=name, min=-mx, max=mx, num_bits=1 + e + mm, is_unsigned=False, is_integer=False,
                    is_packed=False, exponent=e, mantissa=mm)
    m = _FLOAT_RE.match(name)
    if not m:
        raise ValueError(f"unknown storage dtype {name!r}")
    bits, e, mm, kind = int(m.group(1)), int(m.group(2)), int(m.group(3)), m.group(4)
    unsigned = kind == "fnu"
    assert bits == e + mm + (0 if unsigned else 1), name
    bias = 2 ** (e - 1) - 1
    mx = (2.0 - 2.0 ** (-mm)) * 2.0 ** (2 ** e - 1 - bias)  # "fn": every code is a finite number
    return dict(name=name, min=0 if unsigned else -mx, max=mx, num_bits=bits, is_unsigned=unsigned,
                is_integer=False, is_packed=True, exponent=e, mantissa=mm)


# =============================================================================
# bf16 / fp8 value models
# =============================================================================
def bf16_bits(x) -> np.ndarray:
    """float32 -> bfloat16 bit pattern (uint16), round-to-nearest-even, NaN kept quiet."""
    u = np.ascontiguousarray(x, dtype=F32).view(np.uint32)
    r = (u + (((u >> 16) & 1) + np.uint32(0x7FFF))) >> 16
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = np.where(nan, (u >> 16) | 0x0040, r)
    return r.astype(np.uint16)


def from_bf16_bits(b) -> np.ndarray:
    return (np.ascontiguousarray(b).astype(np.uint32) << 16).view(F32)


def bf16_round(x) -> np.ndarray:
    return from_bf16_bits(bf16_bits(x))


_E4M3_POS = None


def _e4m3_table():
    global _E4M3_POS
    if _E4M3_POS is None:
        codes = np.arange(127)  # 0x00..0x7E ; 0x7F is NaN in e4m3fn
        e, m = codes >> 3, codes & 7
        _E4M3_POS = np.where(e == 0, m * 2.0 ** -9, (1 + m / 8.0) * 2.0 ** (e.astype(np.float64) - 7)).astype(F32)
    return _E4M3_POS


def e4m3fn_bits(x) -> np.ndarray:
    """float32 -> float8_e4m3fn bits, RNE, |x| must already be clamped to <= 448 (quant_utils.py:298)."""
    x = np.asarray(x, dtype=F32)
    tab = _e4m3_table()
    ax = np.abs(x)
    lo = np.clip(np.searchsorted(tab, ax, side="right") - 1, 0, 126)
    hi = np.minimum(lo + 1, 126)
    dlo = ax.astype(np.float64) - tab[lo]
    dhi = tab[hi].astype(np.float64) - ax
    pick_hi = (dhi < dlo) | ((dhi == dlo) & (lo % 2 == 1))
    code = np.where(pick_hi, hi, lo).astype(np.uint8)
    code = np.where(np.isnan(x), np.uint8(0x7F), code)
    return (code | (np.signbit(x).astype(np.uint8) << 7)).astype(np.uint8)


def from_e4m3fn_bits(b) -> np.ndarray:
    b = np.asarray(b, dtype=np.uint8)
    mag = np.where((b & 0x7F) == 0x7F, np.nan, _e4m3_table()[np.minimum(b & 0x7F, 126)]).astype(F32)
    return np.where(b & 0x80, -mag, mag).astype(F32)


def e4m3fn_round(x) -> np.ndarray:
    return from_e4m3fn_bits(e4m3fn_bits(x))


def from_e5m2_bits(b) -> np.ndarray:
    """float8_e5m2 bit pattern -> float32 value: an e5m2 code is the upper byte of an IEEE binary16."""
    return (np.asarray(b, dtype=np.uint8).astype(np.uint16) << 8).view(np.float16).astype(F32)


def fma32(a, b, c) -> np.ndarray:
    """float32 fused multiply-add emulated through float64 (a*b is exact in f64)."""
    return (np.asarray(a, F32).astype(np.float64) * np.asarray(b, F32).astype(np.float64)
            + np.asarray(c, F32).astype(np.float64)).astype(F32)


# =============================================================================
# packed integer layouts   (reference: packed_int/pack.py, packed_int/unpack.py)
#
# Each width is described once as a list of bit-runs
#     (word j, first bit in word, value i, first bit in value, run length)
# and both pack and unpack are derived from that table.  Values are taken in
# flattened row-major order, `nv` values per group -> `nw` storage words.
# =============================================================================
def _runs_simple(bits, nv):
    return [(0, bits * i, i, 0, bits) for i in range(nv)]


def _layout(bits):
    if bits in (1, 2, 4):                                   # pack.py:273-321
        nv = 8 // bits
        return 8, nv, 1, _runs_simple(bits, nv)
    if bits == 7:                                           # pack.py:201-221: bit 6-i of v7 rides in bit 7 of byte i
        return 8, 8, 7, [(i, 0, i, 0, 7) for i in range(7)] + [(i, 7, 7, 6 - i, 1) for i in range(7)]
    if bits == 6:                                           # pack.py:225-241
        return 8, 4, 3, [(i, 0, i, 0, 6) for i in range(3)] + [(i, 6, 3, 4 - 2 * i, 2) for i in range(3)]
    if bits == 5:                                           # pack.py:245-269
        r = [(i, 0, i, 0, 5) for i in range(5)]
        r += [(i, 5, 5 + i, 0, 3) for i in range(3)]
        r += [(3, 5, 5, 3, 2), (3, 7, 7, 4, 1), (4, 5, 6, 3, 2), (4, 7, 7, 3, 1)]
        return 8, 8, 5, r
    if bits == 3:                                           # pack.py:280-295
        r = [(i, 0, i, 0, 3) for i in range(3)] + [(i, 3, 3 + i, 0, 3) for i in range(3)]
        r += [(0, 6, 6, 0, 2), (1, 6, 7, 0, 2), (2, 6, 6, 2, 1), (2, 7, 7, 2, 1)]
        return 8, 8, 3, r
    if bits == 15:                                          # pack.py:7-35
        return 16, 16, 15, [(i, 0, i, 0, 15) for i in range(15)] + [(i, 15, 15, 14 - i, 1) for i in range(15)]
    if bits == 14:                                          # pack.py:39-59
        return 16, 8, 7, [(i, 0, i, 0, 14) for i in range(7)] + [(i, 14, 7, 12 - 2 * i, 2) for i in range(7)]
    if bits == 13:                                          # pack.py:63-87
        r = [(i, 0, i, 0, 13) for i in range(13)]
        r += [(3 * j + t, 13, 13 + t, 3 * j, 3) for j in range(4) for t in range(3)]
        r += [(12, 13, 13, 12, 1), (12, 14, 14, 12, 1), (12, 15, 15, 12, 1)]
        return 16, 16, 13, r
    if bits == 12:                                          # pack.py:91-107
        return 16, 4, 3, [(i, 0, i, 0, 12) for i in range(3)] + [(i, 12, 3, 8 - 4 * i, 4) for i in range(3)]
    if bits == 11:                                          # pack.py:111-135
        r = [(i, 0, i, 0, 11) for i in range(8)] + [(i, 11, 8 + i, 0, 5) for i in range(8)]
        r += [(8 + t, 0, 8 + t, 5, 6) for t in range(3)] + [(8 + t, 6, 11 + t, 5, 6) for t in range(3)]
        r += [(8, 12, 14, 5, 4), (9, 12, 15, 5, 4), (10, 12, 14, 9, 2), (10, 14, 15, 9, 2)]
        return 16, 16, 11, r
    if bits == 10:                                          # pack.py:139-163
        r = [(i, 0, i, 0, 10) for i in range(5)] + [(i, 10, 5 + i, 0, 6) for i in range(3)]
        r += [(3, 10, 5, 6, 4), (3, 14, 7, 8, 2), (4, 10, 6, 6, 4), (4, 14, 7, 6, 2)]
        return 16, 8, 5, r
    if bits == 9:                                           # pack.py:167-197
        r = [(i, 0, i, 0, 9) for i in range(8)] + [(i, 9, 8 + i, 0, 7) for i in range(8)]
        r += [(8, 2 * t, 8 + t, 7, 2) for t in range(8)]
        return 16, 16, 9, r
    raise ValueError(f"no packed layout for {bits} bits")


def packed_geometry(bits):
    """(storage word bits, values per group, words per group)."""
    wb, nv, nw, _ = _layout(bits)
    return wb, nv, nw


def pack_uint(codes, bits) -> np.ndarray:
    """unsigned codes (any shape, numel % nv == 0) -> packed words, shape [groups, nw] (1-D for 1/2/4 bit)."""
    wb, nv, nw, runs = _layout(bits)
    v = np.asarray(codes).reshape(-1, nv).astype(np.uint32)
    w = np.zeros((v.shape[0], nw), dtype=np.uint32)
    for (j, wbit, i, vbit, n) in runs:
        w[:, j] |= ((v[:, i] >> vbit) & ((1 << n) - 1)) << wbit
    w = w.astype(np.uint8 if wb == 8 else np.uint16)
    return w.reshape(-1) if nw == 1 else w


def unpack_uint(packed, bits, shape) -> np.ndarray:
    """packed words -> unsigned codes (int32) of `shape`  (unpack.py:6-372)."""
    wb, nv, nw, runs = _layout(bits)
    w = np.asarray(packed)
    if w.dtype.itemsize * 8 != wb:  # uint1 is stored one int64 per packed byte upstream (bool shifts promote)
        w = w.astype(np.int64) & ((1 << wb) - 1)
    w = w.reshape(-1, nw).astype(np.uint32)
    v = np.zeros((w.shape[0], nv), dtype=np.uint32)
    for (j, wbit, i, vbit, n) in runs:
        v[:, i] |= ((w[:, j] >> wbit) & ((1 << n) - 1)) << vbit
    return v.astype(np.int32).reshape(shape)


def pack_int(values, weights_dtype) -> np.ndarray:
    """packed_int/__init__.py:76-80: signed types are stored offset-binary (code = v - min)."""
    info = dtype_info(weights_dtype)
    v = np.asarray(values).astype(np.int64)
    if not info["is_unsigned"]:
        v = v - info["min"]
    return pack_uint(v, info["num_bits"])


def unpack_int(packed, weights_dtype, shape) -> np.ndarray:
    """packed_int/__init__.py:83-88."""
    info = dtype_info(weights_dtype)
    v = unpack_uint(packed, info["num_bits"], shape)
    if not info["is_unsigned"]:
        v = v + info["min"]
    return v


# =============================================================================
# packed minifloats   (reference: packed_float.py:85-132)
# =============================================================================
def decode_minifloat(codes, weights_dtype) -> np.ndarray:
    """eXmY fn/fnu code -> float32 value, by value rather than by bit surgery:
    normal: (1 + m/2^M) * 2^(e - bias); e == 0: subnormal m/2^M * 2^(1 - bias); every code finite.
    (The reference builds the same value by moving fields into an fp32 pattern and then fixing
    subnormals with `2x - sign*min_normal`, packed_float.py:102-131.)"""
    info = dtype_info(weights_dtype)
    E, M, bits = info["exponent"], info["mantissa"], info["num_bits"]
    c = np.asarray(codes).astype(np.int64)
    if info["is_unsigned"]:
        sign = np.zeros_like(c)
        mag = c
    else:
        sign = (c >> (bits - 1)) & 1
        mag = c & ((1 << (bits - 1)) - 1)
    e = mag >> M
    m = mag & ((1 << M) - 1)
    bias = 2 ** (E - 1) - 1
    val = np.where(e == 0, m * 2.0 ** (1 - bias - M), (1.0 + m * 2.0 ** (-M)) * np.exp2((e - bias).astype(np.float64)))
    val = np.where((sign == 1) & (mag != 0), -val, val)   # code "-0" decodes to +0 (packed_float.py:122-123)
    return val.astype(F32)


def unpack_float(packed, weights_dtype, shape) -> np.ndarray:
    info = dtype_info(weights_dtype)
    bits = info["num_bits"]
    if bits in (8, 16):
        codes = np.asarray(packed).astype(np.int64).reshape(shape)
    else:
        codes = unpack_uint(packed, bits, shape)
    return decode_minifloat(codes, weights_dtype)


# =============================================================================
# Hadamard   (reference: quant_utils.py:144-209)
# =============================================================================
def build_hadamard(n: int) -> np.ndarray:
    """float64 matrix: kron^k(H4)/sqrt(n) when n is a power of 4, else Sylvester kron^k(H2)/sqrt(n)."""
    if n & (n - 1) or n < 2:
        raise ValueError(f"Hadamard group size must be a power of 2, got {n}")
    pow4 = (n.bit_length() & 1) == 1
    base = np.array([[1, 1, 1, -1], [1, 1, -1, 1], [1, -1, 1, 1], [-1, 1, 1, 1]], dtype=np.float64) if pow4 \
        else np.array([[1, 1], [1, -1]], dtype=np.float64)
    H = base
    while H.shape[0] < n:
        H = np.kron(H, base)
    return H


def hadamard_matrix(n: int, dtype: str) -> np.ndarray:
    """H / sqrt(n) rounded to `dtype` the way the reference does: divide in dtype (quant_utils.py:150,161)."""
    H = build_hadamard(n).astype(F32)
    Hn = (H / F32(n ** 0.5)).astype(F32)   # div_ by a python scalar is done in f32 opmath
    if dtype == "float16":
        return Hn.astype(np.float16).astype(F32)
    return bf16_round(Hn) if dtype == "bfloat16" else Hn


def rotate_hadamard(x, group: int, dtype: str) -> np.ndarray:
    """x[..., K] -> (x.unflatten(-1,(-1,g)) @ H_g).flatten  computed in `dtype` (fp32 accumulate)."""
    x = np.asarray(x, F32)
    H = hadamard_matrix(group, dtype)
    y = (x.reshape(-1, group) @ H).reshape(x.shape)
    if dtype == "float16":
        return y.astype(np.float16).astype(F32)
    return bf16_round(y) if dtype == "bfloat16" else y.astype(F32)


def hadamard_group_size(channel_size: int, group_size: int):
    """quant_utils.py:212-218."""
    g = min(channel_size, group_size)
    if g & (g - 1):
        g = 2 ** g.bit_length()
    while channel_size % g != 0:
        g //= 2
    return g >= 4, g


# =============================================================================
# per-row quantisers   (reference: quant_utils.py:9-24, 264-299)
# =============================================================================
def quantize_int_mm(w, axis=-1):
    w = np.asarray(w, F32)
    scale = (np.max(np.abs(w), axis=axis, keepdims=True) / F32(127)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.rint((w / scale).astype(F32))
    q = np.clip(q, -128, 127)
    q = np.where(np.isnan(q), 0, q)  # NaN -> integer cast gives 0 on CPU and CUDA
    return q.astype(np.int8), scale


def quantize_uint_mm(w, axis=-1):
    w = np.asarray(w, F32)
    mn = np.min(w, axis=axis, keepdims=True)
    mx = np.max(w, axis=axis, keepdims=True)
    scale = ((mx - mn).astype(F32) / F32(255)).astype(F32)
    zp = (mn - (scale * F32(-128)).astype(F32)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.rint(((w - zp).astype(F32) / scale).astype(F32))
    q = np.clip(q, -128, 127)
    q = np.where(np.isnan(q), 0, q)
    return q.astype(np.int8), scale, zp


def quantize_fp_mm(w, axis=-1):
    """returns (e4m3fn-representable float32 values, scale)."""
    w = np.asarray(w, F32)
    scale = (np.max(np.abs(w), axis=axis, keepdims=True) / F32(448)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (w / scale).astype(F32)
    q = np.nan_to_num(q, nan=0.0, posinf=np.finfo(F32).max, neginf=np.finfo(F32).min)
    q = np.clip(q, -448, 448)
    return e4m3fn_round(q), scale


# =============================================================================
# exact contractions
# =============================================================================
def int_mm(a_i8, b_i8) -> np.ndarray:
    """int8[M,K] @ int8[K,N] -> int32, exact: float32 BLAS on K-chunks of 512 (|partial| < 2^24)."""
    a = np.asarray(a_i8)
    b = np.asarray(b_i8)
    M, K = a.shape
    acc = np.zeros((M, b.shape[1]), dtype=np.int64)
    for k0 in range(0, K, 512):
        acc += np.rint(a[:, k0:k0 + 512].astype(F32) @ b[k0:k0 + 512].astype(F32)).astype(np.int64)
    return acc.astype(np.int32)


def fp8_mm(a, b) -> np.ndarray:
    """fp8-valued float32 operands; every product is exact in f32, sum kept in f64 then rounded once."""
    return (np.asarray(a, np.float64) @ np.asarray(b, np.float64)).astype(F32)


def scaled_mm(acc, sx, sw, bias=None, out_dtype="bfloat16"):
    """kernels/triton_scaled_mm.py:222-232: f32(acc)*sx, then fma(., sw, bias) (or *sw), then cast."""
    t = (np.asarray(acc).astype(F32) * np.asarray(sx, F32)).astype(F32)
    if bias is None:
        y = (t * np.asarray(sw, F32)).astype(F32)
    else:
        y = fma32(t, sw, np.asarray(bias, F32))
    return bf16_round(y) if out_dtype == "bfloat16" else y


# =============================================================================
# weight dequantisation   (reference: dequantizer.py:15-162, 389-429)
# =============================================================================
class Layer:
    """A quantised layer as the reference stores it: logical numpy arrays + the SDNQDequantizer fields."""

    def __init__(self, weight, scale, zero_point=None, svd_up=None, svd_down=None, bias=None, **meta):
        self.weight, self.scale, self.zero_point = weight, scale, zero_point
        self.svd_up, self.svd_down, self.bias = svd_up, svd_down, bias
        self.weights_dtype = meta["weights_dtype"]
        self.quantized_matmul_dtype = meta.get("quantized_matmul_dtype", "int8")
        self.quantized_weight_shape = tuple(meta["quantized_weight_shape"])
        self.result_shape = None if meta.get("result_shape") is None else tuple(meta["result_shape"])
        self.result_dtype = meta.get("result_dtype", "bfloat16")
        self.group_size = meta.get("group_size", -1)
        self.use_quantized_matmul = bool(meta.get("use_quantized_matmul", False))
        self.re_quantize_for_matmul = bool(meta.get("re_quantize_for_matmul", False))
        self.use_hadamard = bool(meta.get("use_hadamard", False))
        self.hadamard_group_size = meta.get("hadamard_group_size", 256)
        self.use_codebook = bool(meta.get("use_codebook", False))
        self.info = dtype_info(self.weights_dtype)
        self.layer_class_name = meta.get("layer_class_name", "Linear")
        self.original_shape = None if meta.get("original_shape") is None else tuple(meta["original_shape"])
        self.svd_dtype = meta.get("svd_dtype", "bfloat16")


def _cast(x, dtype):
    return bf16_round(x) if dtype == "bfloat16" else np.asarray(x, F32)


def unpack_weight(layer: Layer, as_codebook_index=False) -> np.ndarray:
    """dequantizer.py:152-156: packed storage -> integer codes / float values of quantized_weight_shape."""
    info = layer.info
    if not info["is_packed"]:
        return np.asarray(layer.weight)
    if info["is_integer"]:
        return unpack_int(layer.weight, layer.weights_dtype, layer.quantized_weight_shape)
    return unpack_float(layer.weight, layer.weights_dtype, layer.quantized_weight_shape)


def dequantize(layer: Layer, dtype=None, skip_quantized_matmul=False, with_svd=True, non_hadamard=False) -> np.ndarray:
    """SDNQDequantizer.__call__ -> dequantize_weight -> dequantize_{symmetric,asymmetric,codebook}.

    The asymmetric form `addcmul(zp, q, scale)` is a single fused multiply-add in ATen on both CPU (AVX
    fmadd) and CUDA; the golden fixtures are only reproduced bit-exactly with the fused form."""
    dtype = dtype or layer.result_dtype
    q = unpack_weight(layer)
    scale = np.asarray(layer.scale, F32)
    if layer.use_codebook:                                              # dequantizer.py:101-111
        res = np.take_along_axis(scale, q.astype(np.int64), axis=-1) if layer.group_size != -2 else scale[q]
    elif layer.info["is_unsigned"]:                                     # dequantizer.py:27
        zp = np.asarray(layer.zero_point, F32)
        res = fma32(q.astype(F32), scale, zp)
    else:                                                               # dequantizer.py:63
        res = (q.astype(F32) * scale).astype(F32)
    if skip_quantized_matmul and not (layer.re_quantize_for_matmul or layer.info["is_packed"]):
        res = res.T                                                     # dequantizer.py:28-29, 64-65
    if layer.result_shape is not None:
        res = res.reshape(layer.result_shape)
    if with_svd and layer.svd_up is not None:                           # dequantizer.py:32-43
        up, down = np.asarray(layer.svd_up, F32), np.asarray(layer.svd_down, F32)
        if skip_quantized_matmul:
            up, down = up.T, down.T
        if res.ndim > 2 and np.asarray(q).ndim > 2:                     # is_conv (dequantizer.py:30, 36-37): f32 weight + mm in the SVD dtype
            res = (res + _cast((up @ down).astype(F32), layer.svd_dtype).reshape(res.shape)).astype(F32)
        else:
            res = _cast(_cast(res, dtype) + (up @ down).astype(F32), dtype)   # bf16 addmm_: f32 accumulate, one rounding
    res = _cast(res, dtype)
    if layer.use_hadamard and not non_hadamard:                         # dequantizer.py:46-47
        if res.ndim > 2 and np.asarray(q).ndim > 2:                     # is_conv: groups run along the flattened [N, C*kh*kw] weight
            res = rotate_hadamard(res.reshape(res.shape[0], -1), layer.hadamard_group_size, dtype).reshape(res.shape)   # quant_utils.py:199-208
        else:
            res = rotate_hadamard(res, layer.hadamard_group_size, dtype)
    return res


def re_quantize_matmul(layer: Layer):
    """dequantizer.py:204-239 + 166-200: dequant to f32 [N,K] (no SVD, no un-rotate) then row-wise
    re-quantise to the matmul dtype.  Returns logical (Wq[K,N], sw[1,N][, zp[1,N]])."""
    w = dequantize(layer, dtype="float32", with_svd=False, non_hadamard=True)
    if w.ndim > 2:                                                      # convs: flatten(1,-1) (dequantizer.py:167-168, 179-180, 191-192)
        w = w.reshape(w.shape[0], -1)
    mm = dtype_info(layer.quantized_matmul_dtype)
    if mm["is_integer"]:
        if mm["is_unsigned"]:
            q, s, z = quantize_uint_mm(w, axis=-1)
            return q.T, s.T, z.T
        q, s = quantize_int_mm(w, axis=-1)
        return q.T, s.T
    q, s = quantize_fp_mm(w, axis=-1)
    return q.T, s.T


# =============================================================================
# the five Linear forwards   (reference: layers/linear/*.py)
# =============================================================================
def _svd_bias(layer, x_rot, bias, dtype):
    """linear_int8.py:57-62: bias2d = bias + (x @ svd_down) @ svd_up in the SVD dtype."""
    if layer.svd_up is None:
        return bias
    t = _cast(x_rot.astype(F32) @ np.asarray(layer.svd_down, F32), dtype)
    r = (t @ np.asarray(layer.svd_up, F32)).astype(F32)
    if bias is not None:
        r = r + np.asarray(bias, F32)
    return _cast(r, dtype)


def linear_dequant(layer: Layer, x, dtype="bfloat16", skip_quantized_matmul=False):
    """layers/linear/forward.py:24-26 (and the rows<32 guard of every matmul forward)."""
    W = dequantize(layer, dtype=dtype, skip_quantized_matmul=skip_quantized_matmul)
    y = np.asarray(x, F32) @ W.T
    if layer.bias is not None:
        y = y + np.asarray(layer.bias, F32)
    return _cast(y, dtype)


def matmul_inputs(layer: Layer, x, dtype="bfloat16"):
    """get_{int8,uint8,fp8}_matmul_inputs (linear_int8.py:25-71, linear_uint8.py:26-76, linear_fp8.py:25-54).
    Returns dict(xq, wq[K,N], sx[M,1], sw[1,N], bias (None | [N] | [M,N]), x_rot)."""
    mm = dtype_info(layer.quantized_matmul_dtype)
    x = np.asarray(x, F32).reshape(-1, np.asarray(x).shape[-1])
    zp = None
    if layer.re_quantize_for_matmul:
        r = re_quantize_matmul(layer)
        wq, sw = r[0], r[1]
        zp = r[2] if len(r) == 3 else None
    elif layer.info["is_packed"]:
        wq = unpack_weight(layer).T
        sw = np.asarray(layer.scale, F32).T
        zp = None if layer.zero_point is None else np.asarray(layer.zero_point, F32).T
        if not mm["is_integer"]:
            wq = e4m3fn_round(wq)                                       # linear_fp8.py:38
    else:
        wq, sw = np.asarray(layer.weight), np.asarray(layer.scale, F32)
        zp = None if layer.zero_point is None else np.asarray(layer.zero_point, F32)
        if mm["is_integer"] and wq.dtype == np.uint8:                   # linear_int8.py:45-50
            wq = (wq ^ 128).view(np.int8)
            zp = (zp + (sw * F32(128)).astype(F32)).astype(F32) if zp is not None else (sw * F32(128)).astype(F32)
    x_rot = rotate_hadamard(x, layer.hadamard_group_size, dtype) if layer.use_hadamard else x
    bias = _svd_bias(layer, x_rot, layer.bias, dtype)
    K = x.shape[-1]
    if mm["is_integer"] and mm["is_unsigned"]:                          # linear_uint8.py:60-73
        xq, sx, zx = quantize_uint_mm(x_rot, axis=-1)
        colsum = np.sum(np.asarray(wq, np.int64), axis=0, keepdims=True).astype(F32)
        wterm = ((colsum * sw).astype(F32) * zx).astype(F32)
        if zp is not None:
            rowsum = np.sum(xq.astype(np.int64), axis=-1, keepdims=True).astype(F32)
            zb = ((rowsum * sx).astype(F32) * zp).astype(F32)
            zb = (zb + wterm).astype(F32)
            zb = (zb + (F32(K) * (zx * zp).astype(F32)).astype(F32)).astype(F32)
        else:
            zb = wterm
        if bias is not None:
            zb = (zb + np.asarray(bias, F32)).astype(F32)
        bias = zb
    elif mm["is_integer"]:
        xq, sx = quantize_int_mm(x_rot, axis=-1)
        if zp is not None:                                              # linear_int8.py:65-69
            rowsum = np.sum(xq.astype(np.int64), axis=-1, keepdims=True).astype(F32)
            zb = ((rowsum * sx).astype(F32) * zp).astype(F32)
            if bias is not None:
                zb = (zb + np.asarray(bias, F32)).astype(F32)
            bias = zb
    else:
        xq, sx = quantize_fp_mm(x_rot, axis=-1)
    return dict(xq=xq, wq=wq, sx=sx, sw=sw, bias=bias, x_rot=x_rot)


def linear_forward(layer: Layer, x, dtype="bfloat16"):
    """SDNQLinear.forward for any forward_func (forward.py:39-57 dispatch)."""
    x = np.asarray(x, F32)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if not layer.use_quantized_matmul:
        y = linear_dequant(layer, x2, dtype)
    elif x2.shape[0] < 32:                                              # linear_int8.py:102-103
        y = linear_dequant(layer, x2, dtype, skip_quantized_matmul=True)
    else:
        p = matmul_inputs(layer, x2, dtype)
        if dtype_info(layer.quantized_matmul_dtype)["is_integer"]:
            acc = int_mm(p["xq"], p["wq"])
        else:
            acc = fp8_mm(p["xq"], p["wq"])
        y = scaled_mm(acc, p["sx"], p["sw"], p["bias"], out_dtype=dtype)
    return y.reshape(*lead, -1)


# =============================================================================
# convolution forwards   (reference: layers/conv/forward.py, conv_int8.py, conv_uint8.py, conv_fp8.py)
# =============================================================================
def conv_unfold(x, kernel, stride, padding, dilation):
    """F.unfold(x, ...).transpose(1, 2) of process_conv_input (layers/conv/forward.py:30-76) for 2-D convolutions:
    x [B,C,H,W] -> (cols [B, H_out*W_out, C*kh*kw] with columns in (c, i, j) order, (H_out, W_out)).  Zero padding."""
    x = np.asarray(x, F32)
    B, C, H, W = x.shape
    (kh, kw), (sh, sw), (ph, pw), (dh, dw) = kernel, stride, padding, dilation
    Ho = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    Wo = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    xp = np.zeros((B, C, H + 2 * ph, W + 2 * pw), F32)
    xp[:, :, ph:ph + H, pw:pw + W] = x
    cols = np.empty((B, Ho, Wo, C, kh, kw), F32)
    for i in range(kh):
        for j in range(kw):
            cols[:, :, :, :, i, j] = xp[:, :, i * dh:i * dh + sh * (Ho - 1) + 1:sh, j * dw:j * dw + sw * (Wo - 1) + 1:sw].transpose(0, 2, 3, 1)
    return cols.reshape(B, Ho * Wo, C * kh * kw), (Ho, Wo)


def _tuple_n(v, n):
    return (int(v),) * n if isinstance(v, (int, np.integer)) else tuple(int(i) for i in v)


_NP_PAD_MODE = {"reflect": "reflect", "replicate": "edge", "circular": "wrap"}


def conv_forward(layer: Layer, x, kernel_size, stride=1, padding=0, dilation=1, dtype="bfloat16", padding_mode="zeros"):
    """SDNQConv1d/2d.forward for groups == 1 (padding_mode != "zeros": F.pad first, then no padding, layers/conv/forward.py:53-55).  W8A8 layers: conv_{int8,uint8,fp8}_matmul
    (conv_int8.py:17-125) = im2col + the Linear matmul path + permute back; others (and inputs with fewer than 32 rows,
    conv_int8.py:95-96): dequantise + a float convolution (computed here as an f32 im2col GEMM, rounded once)."""
    x = np.asarray(x, F32)
    nd = x.ndim - 2
    k, s, p, d = (_tuple_n(v, nd) for v in (kernel_size, stride, padding, dilation))
    if padding_mode != "zeros":
        x = np.pad(x, [(0, 0), (0, 0)] + [(pi, pi) for pi in p], mode=_NP_PAD_MODE[padding_mode])
        p = (0,) * nd
    x4 = x
    if nd == 1:                                                         # get_conv_args (layers/conv/forward.py:8-27)
        x4 = x[:, :, None, :]
        k, s, p, d = (1, k[0]), (1, s[0]), (0, p[0]), (1, d[0])
    B = x4.shape[0]
    cols, (Ho, Wo) = conv_unfold(x4, k, s, p, d)
    cols2 = cols.reshape(B * Ho * Wo, -1)
    small = x.size / x.shape[2] < 32
    if not layer.use_quantized_matmul or small:
        W = dequantize(layer, dtype=dtype, skip_quantized_matmul=layer.use_quantized_matmul)
        y = cols2 @ W.reshape(W.shape[0], -1).T
        if layer.bias is not None:
            y = y + np.asarray(layer.bias, F32)
        y = _cast(y, dtype)
    else:
        pm = matmul_inputs(layer, cols2, dtype)
        acc = int_mm(pm["xq"], pm["wq"]) if dtype_info(layer.quantized_matmul_dtype)["is_integer"] else fp8_mm(pm["xq"], pm["wq"])
        y = scaled_mm(acc, pm["sx"], pm["sw"], pm["bias"], out_dtype=dtype)
    N = y.shape[-1]
    if nd == 1:
        return y.reshape(B, Wo, N).transpose(0, 2, 1)
    return y.reshape(B, Ho, Wo, N).transpose(0, 3, 1, 2)


# =============================================================================
# fixture loader
# =============================================================================
# ------------------------------------------------------------------------------------------------ quantized attention (row f3)
# Parity status of this block: PINNED.  The reference's attention is a Triton program and cannot run in the authoring container (no
# GPU), so its fixtures were produced on a B200 box: tests/golden/generate_attention.py runs the UNMODIFIED reference (oracle/_ref:
# its quantize_attn and its sdnq_attn_kernel at BLOCK_SIZE_N = 32) and tests/golden/attention_golden.npz holds its inputs, operand
# codes / scales and outputs for nine cases (int8 / e4m3 codes, causal, boolean / additive masks, grouped heads, no smooth-K, head
# dim 128, int8 / e4m3 P.V).  tests/test_oracle_golden.py checks this restatement against them on the CPU (99.8-100 % of the bf16
# outputs identical); tests/test_attention_gpu.py additionally runs the reference kernel live next to the CUDA kernel.

def _cast16(x, dtype):
    return np.asarray(x, F32).astype(np.float16).astype(F32) if dtype == "float16" else _cast(x, dtype)


def quantize_attn(q, k, smooth_k=True, hadamard_group=0, matmul_dtype="int8", dtype="bfloat16"):
    """kernels/triton_atten.py:443-487 for the Q.K^T operands.  q [Z,H,QN,HD], k [Z,KH,KN,HD] (f32 arrays holding `dtype` values)
    -> (q_codes, q_scale [Z,H,QN], k_codes, k_scale [Z,KH,KN]); codes as f32 arrays (ints, or e4m3-representable values)."""
    q = np.asarray(q, dtype=F32)
    k = np.asarray(k, dtype=F32)
    if smooth_k:                                                    # :456-461
        k = (k - k.mean(axis=2, keepdims=True, dtype=F32)).astype(F32)
    if hadamard_group:                                              # :462-466: q in its own dtype, k cast to the rotation's dtype
        q = rotate_hadamard(q, hadamard_group, dtype)
        k = rotate_hadamard(_cast16(k, dtype), hadamard_group, dtype)
    quant = quantize_int_mm if matmul_dtype == "int8" else quantize_fp_mm       # :467-469
    q_q, q_s = quant(q, axis=-1)
    k_q, k_s = quant(k, axis=-1)
    return q_q.astype(F32), q_s.reshape(q.shape[:-1]).astype(F32), k_q.astype(F32), k_s.reshape(k.shape[:-1]).astype(F32)


def quantize_attn_v(v, hadamard_group=0, pv_matmul_dtype="int8", dtype="bfloat16"):
    """kernels/triton_atten.py:478-483: v [Z,VH,KN,HDV] rotated like q / k (in the rotation's dtype) when they are, then quantised per
    key over the head dim -> (v_codes as f32, v_scale [Z,VH,KN])."""
    v = np.asarray(v, dtype=F32)
    if hadamard_group:
        v = rotate_hadamard(_cast16(v, dtype), hadamard_group, dtype)
    quant = quantize_int_mm if pv_matmul_dtype == "int8" else quantize_fp_mm
    v_q, v_s = quant(v, axis=-1)
    return v_q.astype(F32), v_s.reshape(v.shape[:-1]).astype(F32)


def attn_fwd(q_q, k_q, v, q_scale, k_scale, mask=None, is_causal=False, sm_scale=1.0, block_m=128, block_n=128, dtype="bfloat16",
             out_dtype="bfloat16", return_lse=False, v_scale=None, pv_matmul_dtype=None):
    """sdnq_attn_kernel (kernels/triton_atten.py:143-335) with qk_is_quantized=1, use_fp16_accum=0: a block-wise restatement (without
    quantised P.V the result depends on BLOCK_SIZE_N only through rounding; with it, P's row scale is taken per key block, :298-318, so
    block_n is part of the function: 128 is the CUDA kernel's key tile).  q_q [Z,H,QN,HD], k_q [Z,KH,KN,HD] codes, v [Z,VH,KN,HDV]
    values of `dtype`, or with v_scale [Z,VH,KN] the codes of pv_matmul_dtype ("int8" / "float8_e4m3fn"); mask: None, an integer / bool
    array (0 = masked out) or a float array (additive), broadcastable to [Z,H,QN,KN]."""
    q_q, k_q, v = np.asarray(q_q, dtype=F32), np.asarray(k_q, dtype=F32), np.asarray(v, dtype=F32)
    Z, H, QN, _ = q_q.shape
    _, KH, KN, _ = k_q.shape
    _, VH, _, HDV = v.shape
    log2_scale = F32(F32(sm_scale) * F32(1.4426950408889634))      # :189-190
    out = np.zeros((Z, H, QN, HDV), dtype=F32)
    lse = np.zeros((Z, H, QN), dtype=F32)
    if mask is not None:
        mask = np.asarray(mask)
        is_bool = mask.dtype.kind in "biu"
        mask = np.broadcast_to(mask if is_bool else mask.astype(F32), (Z, H, QN, KN))
    ninf = F32(-np.inf)
    with np.errstate(invalid="ignore", over="ignore"):
        for z in range(Z):
            for h in range(H):
                kh, vh = (h * KH) // H, (h * VH) // H               # :197-198
                for m0 in range(0, QN, block_m):
                    m1 = min(QN, m0 + block_m)
                    rows = np.arange(m0, m1)
                    qs = q_scale[z, h, m0:m1, None].astype(F32)
                    m_i = np.full((m1 - m0,), ninf, dtype=F32)      # :231-233
                    l_i = np.ones((m1 - m0,), dtype=F32)
                    acc = np.zeros((m1 - m0, HDV), dtype=F32)
                    for n0 in range(0, KN, block_n):
                        if is_causal and m0 + block_m <= n0:       # :238-239
                            continue
                        n1 = min(KN, n0 + block_n)
                        dot = (q_q[z, h, m0:m1].astype(np.float64) @ k_q[z, kh, n0:n1].astype(np.float64).T).astype(F32)
                        qk = ((dot * qs) * k_scale[z, kh, None, n0:n1].astype(F32)) * log2_scale        # :264-273
                        if is_causal:                               # :277-278
                            qk = np.where(rows[:, None] >= np.arange(n0, n1)[None, :], qk, ninf)
                        if mask is not None:                        # :279-283
                            mk = mask[z, h, m0:m1, n0:n1]
                            qk = np.where(mk != 0, qk, ninf) if is_bool else (qk + mk).astype(F32)
                        m_ij = np.maximum(m_i, qk.max(axis=1))      # :286
                        both = (m_i == ninf) & (m_ij == ninf)       # :287-294 (the do_mask form; identical without a mask)
                        alpha = np.exp2(np.where(both, F32(0), m_i - m_ij)).astype(F32)
                        qk = qk - np.where(m_ij == ninf, F32(0), m_ij)[:, None]
                        pm = np.exp2(qk).astype(F32)                 # :295
                        l_i = fma32(l_i, alpha, pm.sum(axis=1, dtype=F32))       # :296
                        acc = acc * alpha[:, None]                  # :297
                        if v_scale is not None:                     # :298-318
                            pq = (pm * v_scale[z, vh, None, n0:n1].astype(F32)).astype(F32)
                            p_scale = pq.max(axis=1)[:, None].astype(F32)
                            p_scale = (p_scale * F32(1.0 / 127.0 if pv_matmul_dtype == "int8" else 1.0 / 448.0)).astype(F32)
                            p_scale = np.where(p_scale <= F32(2e-38), F32(1), p_scale).astype(F32)
                            inv = (F32(1) / p_scale).astype(F32)
                            if pv_matmul_dtype == "int8":
                                codes = np.floor(fma32(pq, inv, F32(0.5)))
                            else:
                                codes = e4m3fn_round((pq * inv).astype(F32))
                            dot = (codes.astype(np.float64) @ v[z, vh, n0:n1].astype(np.float64)).astype(F32)
                            acc = fma32(dot, p_scale, acc)
                        else:
                            pv = _cast16(pm, dtype).astype(np.float64) @ v[z, vh, n0:n1].astype(np.float64)   # :319-321 (p.to(v.dtype), f32 accumulate)
                            acc = (acc + pv.astype(F32)).astype(F32)
                        m_i = m_ij
                    out[z, h, m0:m1] = acc * (F32(1) / l_i)[:, None]              # :324
                    l = m_i + np.log2(l_i)                                        # :328-330
                    if mask is not None:
                        l = np.where(l == ninf, F32(0), l)
                    lse[z, h, m0:m1] = l
    out = _cast16(out, out_dtype)
    return (out, _cast16(lse, out_dtype)) if return_lse else out


def load_fixture(path):
    """tests/golden/layer_*.npz -> (Layer, arrays dict, meta dict).  `key__T` arrays are the physical
    [N,K] bytes of a K-major tensor and are turned back into the logical transposed view."""
    import json
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    arr = {}
    for k in z.files:
        if k == "meta":
            continue
        a = z[k]
        if k.endswith("__T"):
            arr[k[:-3]] = a.T
        else:
            arr[k] = a
    tinfo = meta["tensors"]

    def get(key):
        if key not in arr:
            return None
        a = arr[key]
        dt = tinfo[key]["dtype"] if key in tinfo and tinfo[key] else None
        if dt == "bfloat16":
            return from_bf16_bits(a)
        if dt == "float8_e4m3fn":
            return from_e4m3fn_bits(a)
        if dt == "float8_e5m2":
            return from_e5m2_bits(a)
        return a

    d = meta["dequantizer"]
    layer = Layer(get("weight"), get("scale"), get("zero_point"), get("svd_up"), get("svd_down"),
                  bias=from_bf16_bits(arr["bias"]) if "bias" in arr else None, **d)
    return layer, arr, meta
