"""Load-time packers for sub-byte integers and eXmY minifloats (torch, any device) and the public unpack entry points.

`pack_*` run once per layer at quantisation time, so they stay in PyTorch; the bit layouts they emit are the ones the CUDA
unpack kernels (csrc/unpack.cuh) read and are pinned against reference-generated vectors (tests/test_host_api.py).
Layout source: reference packed_int/pack.py:6-321 and packed_float.py:26-82, restated as one table of bit runs
    (storage word, first bit in word, value index, first bit in value, run length)
per width (the same table the oracle uses, written independently here for torch).
`unpack_int` / `unpack_float` dispatch to the CUDA kernel for device tensors; the torch table path below is only used for
CPU tensors at load time (e.g. dequantising a layer that is still on the host) and is not on the forward path."""
import torch

from .common import dtype_dict


def _bit_runs(bits: int):
    """-> (word_bits, values per group, words per group, runs)."""
    if bits in (1, 2, 4):
        nv = 8 // bits
        return 8, nv, 1, [(0, bits * i, i, 0, bits) for i in range(nv)]
    if bits == 3:
        runs = [(i, 0, i, 0, 3) for i in range(3)] + [(i, 3, 3 + i, 0, 3) for i in range(3)]
        return 8, 8, 3, runs + [(0, 6, 6, 0, 2), (1, 6, 7, 0, 2), (2, 6, 6, 2, 1), (2, 7, 7, 2, 1)]
    if bits == 5:
        runs = [(i, 0, i, 0, 5) for i in range(5)] + [(i, 5, 5 + i, 0, 3) for i in range(3)]
        return 8, 8, 5, runs + [(3, 5, 5, 3, 2), (3, 7, 7, 4, 1), (4, 5, 6, 3, 2), (4, 7, 7, 3, 1)]
    if bits == 6:
        return 8, 4, 3, [(i, 0, i, 0, 6) for i in range(3)] + [(i, 6, 3, 4 - 2 * i, 2) for i in range(3)]
    if bits == 7:
        return 8, 8, 7, [(i, 0, i, 0, 7) for i in range(7)] + [(i, 7, 7, 6 - i, 1) for i in range(7)]
    if bits == 9:
        runs = [(i, 0, i, 0, 9) for i in range(8)] + [(i, 9, 8 + i, 0, 7) for i in range(8)]
        return 16, 16, 9, runs + [(8, 2 * t, 8 + t, 7, 2) for t in range(8)]
    if bits == 10:
        runs = [(i, 0, i, 0, 10) for i in range(5)] + [(i, 10, 5 + i, 0, 6) for i in range(3)]
        return 16, 8, 5, runs + [(3, 10, 5, 6, 4), (3, 14, 7, 8, 2), (4, 10, 6, 6, 4), (4, 14, 7, 6, 2)]
    if bits == 11:
        runs = [(i, 0, i, 0, 11) for i in range(8)] + [(i, 11, 8 + i, 0, 5) for i in range(8)]
        runs += [(8 + t, 0, 8 + t, 5, 6) for t in range(3)] + [(8 + t, 6, 11 + t, 5, 6) for t in range(3)]
        return 16, 16, 11, runs + [(8, 12, 14, 5, 4), (9, 12, 15, 5, 4), (10, 12, 14, 9, 2), (10, 14, 15, 9, 2)]
    if bits == 12:
        return 16, 4, 3, [(i, 0, i, 0, 12) for i in range(3)] + [(i, 12, 3, 8 - 4 * i, 4) for i in range(3)]
    if bits == 13:
        runs = [(i, 0, i, 0, 13) for i in range(13)] + [(3 * j + t, 13, 13 + t, 3 * j, 3) for j in range(4) for t in range(3)]
        return 16, 16, 13, runs + [(12, 13, 13, 12, 1), (12, 14, 14, 12, 1), (12, 15, 15, 12, 1)]
    if bits == 14:
        return 16, 8, 7, [(i, 0, i, 0, 14) for i in range(7)] + [(i, 14, 7, 12 - 2 * i, 2) for i in range(7)]
    if bits == 15:
        return 16, 16, 15, [(i, 0, i, 0, 15) for i in range(15)] + [(i, 15, 15, 14 - i, 1) for i in range(15)]
    raise ValueError(f"no packed layout for {bits}-bit values")


def _pack_codes(codes: torch.Tensor, bits: int) -> torch.Tensor:
    word_bits, nv, nw, runs = _bit_runs(bits)
    v = codes.contiguous().reshape(-1, nv).to(torch.int32)
    words = torch.zeros((v.shape[0], nw), dtype=torch.int32, device=v.device)
    for word, wbit, val, vbit, n in runs:
        words[:, word] |= ((v[:, val] >> vbit) & ((1 << n) - 1)) << wbit
    if word_bits == 8:
        out = words.to(torch.uint8)
    else:
        out = torch.where(words >= 32768, words - 65536, words).to(torch.int16)      # two's complement int16 storage
    if bits == 1:
        out = out.to(torch.int64)       # upstream stores uint1 as one int64 per packed byte; keep checkpoints interchangeable
    return out.reshape(-1) if nw == 1 else out


def _unpack_codes(packed: torch.Tensor, bits: int, shape) -> torch.Tensor:
    word_bits, nv, nw, runs = _bit_runs(bits)
    w = packed.reshape(-1, nw).to(torch.int32) & ((1 << word_bits) - 1)
    v = torch.zeros((w.shape[0], nv), dtype=torch.int32, device=w.device)
    for word, wbit, val, vbit, n in runs:
        v[:, val] |= ((w[:, word] >> wbit) & ((1 << n) - 1)) << vbit
    return v.reshape(tuple(shape))


def pack_int(tensor: torch.Tensor, weights_dtype: str) -> torch.Tensor:
    """reference packed_int/__init__.py:76-80 (signed types are stored offset-binary)."""
    info = dtype_dict[weights_dtype]
    codes = tensor.to(torch.int32)
    if not info["is_unsigned"]:
        codes = codes - info["min"]
    return _pack_codes(codes, info["num_bits"])


def unpack_int(packed_tensor: torch.Tensor, weights_dtype: str, shape, dtype: torch.dtype | None = None) -> torch.Tensor:
    """reference packed_int/__init__.py:83-88.  CUDA tensors go through the unpack kernel."""
    info = dtype_dict[weights_dtype]
    if packed_tensor.is_cuda and info["num_bits"] <= 8:
        from . import ops
        if dtype is None or info["is_unsigned"]:
            dtype = torch.uint8 if info["is_unsigned"] else info["torch_dtype"]
        return ops.unpack(packed_tensor, weights_dtype, shape, dtype=dtype)
    codes = _unpack_codes(packed_tensor, info["num_bits"], shape)
    if info["is_unsigned"]:
        return codes.to(torch.uint8 if info["num_bits"] <= 8 else torch.int16)
    return (codes + info["min"]).to(info["torch_dtype"] if dtype is None else dtype)


def _minifloat_fields(weights_dtype: str):
    info = dtype_dict[weights_dtype]
    return info["num_bits"], info["exponent"], info["mantissa"], info["is_unsigned"]


def encode_minifloat(x: torch.Tensor, weights_dtype: str) -> torch.Tensor:
    """float values (already clamped to the format's range) -> integer codes.

    Rounding follows the reference encoder bit-for-bit (packed_float.py:26-82), which is *not* plain round-to-nearest-even:
    normals round the dropped mantissa bits up when  (dropped & ~0b1111-low-bits) > half ; subnormals are computed as
    round(|x| * 2^M / min_normal) (round-half-even on the scaled value)."""
    bits, E, M, unsigned = _minifloat_fields(weights_dtype)
    drop = 23 - M
    xi = x.to(torch.float32).contiguous().view(torch.int32)
    # only the top four dropped bits [drop-4, drop) take part in the comparison, and it is a strict ">" against one half
    # (every format has M <= 15, so drop >= 8)
    top4 = (-(1 << (drop - 4))) & ~(-(1 << drop))
    round_up = (xi & top4) > (1 << (drop - 1))
    xi = torch.where(round_up, xi + (1 << drop), xi)
    if E < 8:
        min_normal = 2.0 ** (2 - (1 << (E - 1)))
        ax = xi.view(torch.float32).abs()
        sub = ax < min_normal
        sub_bits = (xi & -2147483648) | ((ax * ((1 << M) / min_normal)).round().to(torch.int32) << drop)
        xi = torch.where(sub, sub_bits, xi)
    xi = xi >> drop
    sign_mask = (1 << (bits - 1)) if unsigned else (1 << (bits - 1)) + (1 << (bits - 2))
    code = (((xi >> (8 - E)) & sign_mask) | (xi & ~sign_mask)) & ((1 << bits) - 1)
    return code


def decode_minifloat(codes: torch.Tensor, weights_dtype: str) -> torch.Tensor:
    """integer codes -> float32 values: (1+m/2^M)*2^(e-bias), subnormals m/2^M*2^(1-bias), "-0" -> +0 (packed_float.py:85-132)."""
    bits, E, M, unsigned = _minifloat_fields(weights_dtype)
    c = codes.to(torch.int32)
    magbits = bits if unsigned else bits - 1
    mag = c & ((1 << magbits) - 1)
    e, m = mag >> M, mag & ((1 << M) - 1)
    bias = (1 << (E - 1)) - 1
    normal = (1.0 + m.to(torch.float32) * 2.0 ** (-M)) * torch.exp2((e - bias).to(torch.float32))
    sub = m.to(torch.float32) * 2.0 ** (1 - bias - M)
    val = torch.where(e == 0, sub, normal)
    if not unsigned:
        neg = ((c >> (bits - 1)) & 1).bool() & (mag != 0)
        val = torch.where(neg, -val, val)
    return val


def pack_float(x: torch.Tensor, weights_dtype: str) -> torch.Tensor:
    bits = dtype_dict[weights_dtype]["num_bits"]
    codes = encode_minifloat(x, weights_dtype)
    if bits in (8, 16):
        return codes.to(torch.uint8) if bits == 8 else torch.where(codes >= 32768, codes - 65536, codes).to(torch.int16).view(torch.uint16)
    return _pack_codes(codes, bits)


def unpack_float(x: torch.Tensor, weights_dtype: str, shape) -> torch.Tensor:
    bits = dtype_dict[weights_dtype]["num_bits"]
    if x.is_cuda and bits <= 8:
        from . import ops
        return ops.unpack(x, weights_dtype, shape, dtype=torch.float32)
    codes = x.to(torch.int32).reshape(tuple(shape)) & ((1 << bits) - 1) if bits in (8, 16) else _unpack_codes(x, bits, shape)
    return decode_minifloat(codes, weights_dtype)
