"""The per-lane arithmetic of the CUDA kernels, compiled for the host and checked against the oracle WITHOUT a GPU.

`tests/host_emu/emu.cpp` includes `sdnq_b200/csrc/unpack.cuh` (and through it `common.cuh`) unchanged and is built with g++
(together with the kernels' `.cu` files, see `tests/host_emu/build_emu.py` and `tests/test_emulated_gpu_suite.py`);
`tests/host_emu/prelude.h` supplies host stand-ins for the few device intrinsics those headers use (`__byte_perm`,
`__uint_as_float`, ...).  So the storage decoders (`load_octet_bytes` + `decode_octet`), the value decoders
(`codes_to_values`: signed offset, every minifloat format, fp8) and the byte-permute fast path (`octet_to_floats`) that the
dequant / re-quantise / GEMV kernels run are the same source lines the GPU executes; the bar is bit-exact against the numpy oracle,
which is itself pinned to reference-generated fixtures (tests/test_oracle_golden.py).  The GPU tests stay the parity tests
proper; this is the CPU-only regression gate for the integer / bit-manipulation part of the kernels."""
import ctypes
import os
import zlib

import numpy as np
import pytest

from oracle import sdnq_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "host_emu")
P = ctypes.c_void_p


class WeightFormat(ctypes.Structure):            # include/sdnq_b200.h: sdnq_weight_format
    _fields_ = [(n, ctypes.c_int32) for n in ("kind", "bits", "is_unsigned", "exponent", "mantissa", "word_bytes")]


W_INT, W_MINIFLOAT, W_E4M3, W_E5M2 = range(4)


@pytest.fixture(scope="module")
def emu():
    from tests.host_emu import build_emu
    out = build_emu.build()              # g++ build of the kernels' CUDA sources + tests/host_emu/emu.cpp (cached by content hash)
    lib = ctypes.CDLL(out)
    lib.emu_decode.argtypes = [ctypes.c_int, ctypes.c_int, P, ctypes.c_int64, P]
    lib.emu_values.argtypes = [ctypes.POINTER(WeightFormat), P, ctypes.c_int64, P]
    lib.emu_octet_to_floats.argtypes = [ctypes.POINTER(WeightFormat), P, ctypes.c_int64, P]
    for name in ("emu_f32_to_e4m3", "emu_e4m3_to_f32", "emu_e5m2_to_f32", "emu_round_bf16", "emu_round_f16"):
        getattr(lib, name).argtypes = [P, ctypes.c_int64, P]
        getattr(lib, name).restype = None
    lib.emu_hadamard_sign.restype = ctypes.c_uint32
    return lib


def _ptr(a):
    return a.ctypes.data_as(P)


def _storage_bytes(packed, bits):
    """the oracle's packed words as the byte stream the kernels read (uint1: upstream keeps one int64 per packed byte)"""
    a = np.ascontiguousarray(packed)
    assert a.dtype == np.uint8, a.dtype
    return a.reshape(-1)


INT_NAMES = [f"{u}int{b}" for u in ("", "u") for b in range(2, 9)] + ["uint1"]
_FLOATS = [f"float{1 + e + m}_e{e}m{m}fn" for e in range(1, 6) for m in range(0, 7) if 2 <= 1 + e + m <= 8] + \
          [f"float{e + m}_e{e}m{m}fnu" for e in range(1, 6) for m in range(0, 8) if 1 <= e + m <= 8]


def _known_minifloats():
    names = []
    for n in _FLOATS + ["float8_e4m3fn_sdnq"]:
        try:
            info = O.dtype_info(n)
        except (ValueError, AssertionError):
            continue
        if info["is_packed"] or n.endswith("_sdnq"):
            names.append(n)
    return names


MINIFLOATS = _known_minifloats()


def test_the_format_lists_cover_the_dtype_table():
    from sdnq_b200.common import dtype_dict
    canonical = {k for k, v in dtype_dict.items()
                 if k.startswith("float") and not v["is_integer"] and v["is_packed"] and v["num_bits"] <= 8}
    assert canonical <= set(MINIFLOATS), sorted(canonical - set(MINIFLOATS))
    assert len(MINIFLOATS) >= 55


@pytest.mark.parametrize("name", INT_NAMES)
def test_integer_storage_decoders_bit_exact(emu, name):
    info = O.dtype_info(name)
    bits = info["num_bits"]
    rng = np.random.default_rng(bits * 7 + info["is_unsigned"])
    n = 8 * 1031                                                   # octets at every byte alignment for the odd widths
    values = rng.integers(info["min"], info["max"] + 1, size=n)
    values[:2 ** min(bits, 8)] = np.arange(info["min"], info["min"] + 2 ** min(bits, 8))[:n]       # every code at least once
    packed = O.pack_int(values, name) if bits < 8 else (values.astype(np.int8).view(np.uint8) if not info["is_unsigned"] else values.astype(np.uint8))
    raw = _storage_bytes(packed, bits)
    assert raw.size == n * bits // 8
    # unsigned codes
    codes = np.empty(n, dtype=np.uint32)
    assert emu.emu_decode(bits, 1, _ptr(raw), n // 8, _ptr(codes)) == 0
    want_codes = values - info["min"] if (bits < 8 and not info["is_unsigned"]) else (values & 0xFF)
    np.testing.assert_array_equal(codes.astype(np.int64), want_codes)
    # real values through both value paths
    fmt = WeightFormat(W_INT, bits, int(info["is_unsigned"]), 0, 0, 1)
    for fn in (emu.emu_values, emu.emu_octet_to_floats):
        out = np.empty(n, dtype=np.float32)
        assert fn(ctypes.byref(fmt), _ptr(raw), n // 8, _ptr(out)) == 0
        np.testing.assert_array_equal(out, values.astype(np.float32))
    if bits < 8:
        np.testing.assert_array_equal(O.unpack_int(packed, name, (n,)), values)       # and the oracle agrees with itself


def test_uint1_int64_words(emu):
    """upstream stores uint1 as one int64 per packed byte; the kernels read the low byte of each word"""
    rng = np.random.default_rng(1)
    bits_in = rng.integers(0, 2, size=8 * 257)
    packed = O.pack_int(bits_in, "uint1").astype(np.int64)
    raw = packed.view(np.uint8)
    codes = np.empty(bits_in.size, dtype=np.uint32)
    assert emu.emu_decode(1, 8, _ptr(raw), bits_in.size // 8, _ptr(codes)) == 0
    np.testing.assert_array_equal(codes, bits_in)


@pytest.mark.parametrize("name", MINIFLOATS)
def test_minifloat_decoders_bit_exact(emu, name):
    info = O.dtype_info(name)
    bits = info["num_bits"]
    rng = np.random.default_rng(bits)
    n = 8 * 257
    codes = rng.integers(0, 2 ** bits, size=n)
    codes[:2 ** bits] = np.arange(2 ** bits)[:n]                   # the whole code table
    packed = O.pack_uint(codes, bits) if bits < 8 else codes.astype(np.uint8)
    raw = _storage_bytes(packed, bits)
    fmt = WeightFormat(W_MINIFLOAT, bits, int(info["is_unsigned"]), info["exponent"], info["mantissa"], 1)
    out = np.empty(n, dtype=np.float32)
    assert emu.emu_values(ctypes.byref(fmt), _ptr(raw), n // 8, _ptr(out)) == 0
    want = O.unpack_float(packed, name, (n,))
    np.testing.assert_array_equal(out.view(np.uint32), want.view(np.uint32))         # bit patterns: "-0" must decode to +0


def test_native_fp8_tables(emu):
    allb = np.arange(256, dtype=np.uint8)
    out = np.empty(256, dtype=np.float32)
    emu.emu_e4m3_to_f32(_ptr(allb), 256, _ptr(out))
    want = O.from_e4m3fn_bits(allb)
    finite = ~np.isnan(want)
    np.testing.assert_array_equal(out[finite], want[finite])
    assert np.isnan(out[~finite]).all() and set(allb[~finite]) == {0x7F, 0xFF}
    # e5m2 is the top byte of an IEEE half
    emu.emu_e5m2_to_f32(_ptr(allb), 256, _ptr(out))
    want5 = (allb.astype(np.uint16) << 8).view(np.float16).astype(np.float32)
    ok = np.isfinite(want5)
    np.testing.assert_array_equal(out[ok], want5[ok])
    # through the weight-format path as well
    for kind, table in ((W_E4M3, want), (W_E5M2, want5)):
        fmt = WeightFormat(kind, 8, 0, 4 if kind == W_E4M3 else 5, 3 if kind == W_E4M3 else 2, 1)
        assert emu.emu_values(ctypes.byref(fmt), _ptr(allb), 32, _ptr(out)) == 0
        good = np.isfinite(table)
        np.testing.assert_array_equal(out[good], table[good])


def test_activation_cast_to_e4m3_is_round_to_nearest_even(emu):
    """K2's fp8 quantiser: clamp to +-448 then cast (quant_utils.py:289-299); the cast must equal torch's / the oracle's RNE"""
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(20000).astype(np.float32) * s for s in (1e-3, 0.05, 1.0, 30.0, 200.0)])
    table = O.from_e4m3fn_bits(np.arange(0x7F, dtype=np.uint8))                        # positive finite values, ascending
    mids = (table[:-1] + table[1:]) / 2                                                  # exact ties between neighbours
    x = np.clip(np.concatenate([x, mids, -mids, table, -table, np.float32([0.0, -0.0, 448.0, -448.0, 2 ** -10, 2 ** -11])]), -448, 448).astype(np.float32)
    got = np.empty(x.size, dtype=np.uint8)
    emu.emu_f32_to_e4m3(_ptr(x), x.size, _ptr(got))
    np.testing.assert_array_equal(got, O.e4m3fn_bits(x))


def test_rounding_to_the_activation_dtype(emu):
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(50000) * np.exp2(rng.integers(-20, 20, size=50000))).astype(np.float32)
    out = np.empty_like(x)
    emu.emu_round_bf16(_ptr(x), x.size, _ptr(out))
    np.testing.assert_array_equal(out, O.bf16_round(x))
    emu.emu_round_f16(_ptr(x), x.size, _ptr(out))
    with np.errstate(over="ignore"):
        np.testing.assert_array_equal(out, x.astype(np.float16).astype(np.float32))


@pytest.mark.parametrize("G", [4, 16, 64, 256])
def test_h4_family_sign_and_placement_rules(emu, G):
    """common.cuh factors the reference's power-of-4 Hadamard (quant_utils.py:155-165, kron^k(H4)) as
    P . D . Sylvester . D: D negates positions with a base-4 digit equal to 3 (product over digits), P swaps the two bits of
    every base-4 digit of the output index.  Check the sign rule and the placement map the kernels use against the matrix."""
    H = O.build_hadamard(G).astype(np.float64)                        # unnormalised +-1 entries
    H = np.sign(H)
    logg = G.bit_length() - 1
    S = np.array([[1.0]])
    for _ in range(logg):
        S = np.kron(S, np.array([[1.0, 1.0], [1.0, -1.0]]))
    pos = np.arange(G)
    sign = np.array([-1.0 if (emu.emu_hadamard_sign(G, int(p) // 8, int(p) % 8) >> 31) else 1.0 for p in pos])
    digits3 = np.ones(G)
    for d in range(logg // 2):
        digits3 *= np.where(((pos >> (2 * d)) & 3) == 3, -1.0, 1.0)
    np.testing.assert_array_equal(sign, digits3)
    # output p of D.S.D lands at swap_bit_pairs(p): bits (0,1) are exchanged in registers by hadamard_warp, the rest by hadamard_dest
    swap = np.zeros(G, dtype=np.int64)
    for p in pos:
        q = 0
        for d in range(logg // 2):
            dig = (p >> (2 * d)) & 3
            q |= (((dig & 1) << 1) | (dig >> 1)) << (2 * d)
        swap[p] = q
    M = np.zeros((G, G))
    M[swap, :] = (sign[:, None] * S * sign[None, :])
    np.testing.assert_array_equal(M, H)
    for p in range(0, G, 4):
        lane, half = p // 8, (p % 8) // 4
        dest = emu.emu_hadamard_dest(G, lane, half)
        assert dest == (swap[p] & ~3) and dest % 4 == 0, (p, dest, swap[p])


# ---- warp-level code on 32 lock-stepped host threads (tests/host_emu/warp.h: host models of mma.sync m16n8k16, movmatrix, cvt pack,
# shfl.xor following the PTX fragment layouts).  Inputs are small integers times a power of two, so every partial sum is exact in
# f32 and the tensor-core / butterfly / reference orders of summation must agree bit for bit.
def _exact_chunks(rng, chunks):
    x = rng.integers(-64, 65, size=(chunks, 256)).astype(np.float32) * np.exp2(rng.integers(-3, 4, size=(chunks, 1))).astype(np.float32)
    x[0, :] = 0.0
    x[0, 5] = 1.0                                                    # a unit impulse: one column of the matrix
    return x


def _bind_rotations(emu):
    emu.emu_rotate_tc.argtypes = [ctypes.c_int, ctypes.c_int, P, P, P, ctypes.c_int64]
    emu.emu_rotate_tc.restype = None
    emu.emu_rotate_butterfly.argtypes = [ctypes.c_int, ctypes.c_int, P, P, ctypes.c_int64]
    emu.emu_rotate_butterfly.restype = None


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
@pytest.mark.parametrize("G", [4, 8, 16, 32, 64, 128, 256])
def test_tensor_core_rotation_on_a_host_warp(emu, G, dtype):
    """hadtc::Rotation<T>::apply -- the default rotation of K2 (FLUX: fp8 + Hadamard-256) and of the rotated K3 -- executed from
    the kernel source with its compile-time fragment tables, against the reference rotation (quant_utils.py:193-209)."""
    _bind_rotations(emu)
    f16 = dtype == "float16"
    rng = np.random.default_rng(G + f16)
    chunks = 6
    x = _exact_chunks(rng, chunks)
    bits = (x.astype(np.float16).view(np.uint16) if f16 else O.bf16_bits(x).astype(np.uint16)).reshape(-1)
    out = np.empty_like(bits)
    lane_max = np.empty(32 * chunks, dtype=np.float32)
    emu.emu_rotate_tc(int(f16), G, _ptr(bits), _ptr(out), _ptr(lane_max), chunks)
    got = (out.view(np.float16).astype(np.float32) if f16 else O.from_bf16_bits(out)).reshape(chunks, 256)
    want = O.rotate_hadamard(x, G, dtype)
    np.testing.assert_array_equal(got, want)
    # the value each lane returns: max |rotated| over its eight outputs before the rounding to T
    exact = (x.reshape(-1, G).astype(np.float64) @ O.hadamard_matrix(G, dtype).astype(np.float64)).reshape(chunks, 256)
    own = np.concatenate([exact[:, :128].reshape(chunks, 32, 4), exact[:, 128:].reshape(chunks, 32, 4)], axis=2)
    np.testing.assert_array_equal(lane_max.reshape(chunks, 32), np.abs(own).max(axis=2).astype(np.float32))


@pytest.mark.parametrize("G", [4, 8, 16, 32, 64, 128, 256])
def test_butterfly_rotation_on_a_host_warp(emu, G):
    """hadamard_warp<G> + hadamard_dest<G>: the shuffle-butterfly rotation (f32 activations, the gather conv quantiser, A/B knob)"""
    _bind_rotations(emu)
    rng = np.random.default_rng(100 + G)
    chunks = 4
    x = _exact_chunks(rng, chunks)
    out = np.empty_like(x)
    emu.emu_rotate_butterfly(0, G, _ptr(np.ascontiguousarray(x)), _ptr(out), chunks)
    np.testing.assert_array_equal(O.bf16_round(out), O.rotate_hadamard(x, G, "bfloat16"))


def test_tensor_core_rotation_on_random_activations(emu):
    """random bf16 activations: sums are no longer exact, so allow the one-ulp ties the GPU test allows (tests/test_kernels_gpu.py)"""
    _bind_rotations(emu)
    rng = np.random.default_rng(7)
    chunks = 16
    x = O.bf16_round(rng.standard_normal((chunks, 256)).astype(np.float32))
    bits = O.bf16_bits(x).astype(np.uint16).reshape(-1)
    out = np.empty_like(bits)
    lane_max = np.empty(32 * chunks, dtype=np.float32)
    emu.emu_rotate_tc(0, 256, _ptr(bits), _ptr(out), _ptr(lane_max), chunks)
    got = O.from_bf16_bits(out).reshape(chunks, 256)
    want = O.rotate_hadamard(x, 256, "bfloat16")
    ulp = np.exp2(np.floor(np.log2(np.maximum(np.abs(want), 1e-30))) - 7)
    # one bf16 ulp of the result, or -- where the 256 terms cancel to almost nothing -- the f32 accumulation noise of sums of size ~16
    assert (np.abs(got - want) <= np.maximum(ulp, 4e-6)).all()
    assert (got != want).mean() < 0.01


# ---- a whole kernel on emulated CTAs: K5p, the small-M Linear from the stored (packed / grouped) weight (gemv_packed_kernel.cuh).
# 256 lock-stepped host threads per CTA run the kernel body unchanged; loads, fragment mapping, K-split reduction through
# "shared memory", tails in N / K / M and the epilogue are all the device source.
GEMV_CASES = [
    # weights_dtype, group (0 = row-wise), M, N, K, bias kind, activation dtype
    ("int4", 0, 4, 40, 208, "vector", "bfloat16"),
    ("int4", 128, 1, 16, 256, "none", "bfloat16"),
    ("uint4", 64, 9, 24, 192, "matrix", "bfloat16"),
    ("int2", 16, 32, 33, 64, "vector", "bfloat16"),
    ("uint3", 32, 5, 17, 96, "vector", "float16"),
    ("int5", 0, 2, 8, 80, "none", "bfloat16"),
    ("uint6", 16, 17, 16, 48, "vector", "bfloat16"),
    ("int7", 8, 3, 19, 16, "vector", "bfloat16"),
    ("int8", 0, 4, 48, 1040, "vector", "bfloat16"),
    ("uint8", 0, 6, 20, 128, "matrix", "float16"),
    ("float6_e3m2fn", 32, 4, 16, 128, "vector", "bfloat16"),
    ("float4_e2m1fn", 0, 8, 31, 64, "none", "float16"),
    ("float8_e4m3fn", 0, 4, 32, 576, "vector", "bfloat16"),
]


def _stored_weight(rng, name, N, K, exact):
    """random codes of the format -> (storage bytes as the reference keeps them, real values q[N,K] before the scale, WeightFormat)"""
    info = O.dtype_info(name)
    bits = info["num_bits"]
    if info["is_integer"]:
        vals = rng.integers(info["min"], info["max"] + 1, size=(N, K))
        if bits == 8:
            raw = vals.astype(np.uint8) if info["is_unsigned"] else vals.astype(np.int8).view(np.uint8)
        else:
            raw = O.pack_int(vals, name)
        return np.ascontiguousarray(raw).reshape(-1), vals.astype(np.float32), WeightFormat(W_INT, bits, int(info["is_unsigned"]), 0, 0, 1)
    if name == "float8_e4m3fn":
        codes = rng.integers(0, 256, size=(N, K)).astype(np.uint8)
        codes[(codes & 0x7F) == 0x7F] = 0x3C                                      # no NaN codes
        return codes.reshape(-1), O.from_e4m3fn_bits(codes), WeightFormat(W_E4M3, 8, 0, 4, 3, 1)
    codes = rng.integers(0, 2 ** bits, size=(N, K))
    raw = O.pack_uint(codes, bits) if bits < 8 else codes.astype(np.uint8)
    fmt = WeightFormat(W_MINIFLOAT, bits, int(info["is_unsigned"]), info["exponent"], info["mantissa"], 1)
    return np.ascontiguousarray(raw).reshape(-1), O.decode_minifloat(codes, name), fmt


def _round_to(x, dtype):
    return O.bf16_round(x) if dtype == "bfloat16" else np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


def _bits_of(x, dtype):
    return O.bf16_bits(x).astype(np.uint16) if dtype == "bfloat16" else np.asarray(x, np.float32).astype(np.float16).view(np.uint16)


def _from_bits(b, dtype):
    return O.from_bf16_bits(b) if dtype == "bfloat16" else b.view(np.float16).astype(np.float32)


@pytest.mark.parametrize("exact", [True, False], ids=["exact_sums", "random"])
@pytest.mark.parametrize("case", GEMV_CASES, ids=[f"{c[0]}-g{c[1]}-M{c[2]}" for c in GEMV_CASES])
def test_small_m_packed_linear_kernel_on_emulated_ctas(emu, case, exact):
    name, group, M, N, K, bias_kind, dtype = case
    emu.emu_gemv_packed.argtypes = [P, ctypes.c_int, ctypes.c_int64, P, ctypes.POINTER(WeightFormat), P, P, ctypes.c_int64, P, ctypes.c_int,
                                    ctypes.c_int64, P, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
    info = O.dtype_info(name)
    rng = np.random.default_rng(zlib.crc32(repr((name, group, M, exact)).encode()))
    raw, q, fmt = _stored_weight(rng, name, N, K, exact)
    raw = np.concatenate([raw, np.zeros(16, np.uint8)])                            # the kernel never reads past the weight; slack only for alignment of the copy
    g = group or K
    if exact:       # power-of-two scales and small integer activations: every product and partial sum is exact in f32
        scale = np.exp2(rng.integers(-4, 2, size=(N, K // g))).astype(np.float32)
        zp = np.exp2(rng.integers(-3, 1, size=(N, K // g))).astype(np.float32) * rng.integers(-4, 5, size=(N, K // g)) if info["is_unsigned"] else None
        x = rng.integers(-8, 9, size=(M, K)).astype(np.float32)
        bias_vals = rng.integers(-16, 17, size=(M, N)).astype(np.float32)
    else:
        scale = (rng.random((N, K // g)).astype(np.float32) + 0.5) * np.float32(0.02)
        zp = (rng.standard_normal((N, K // g)).astype(np.float32) * np.float32(0.1)) if info["is_unsigned"] else None
        x = _round_to(rng.standard_normal((M, K)).astype(np.float32), dtype)
        bias_vals = _round_to(rng.standard_normal((M, N)).astype(np.float32), dtype)
    if zp is not None:
        zp = np.ascontiguousarray(zp, dtype=np.float32)
    # the reference's dequantised weight: f32 math in its order, rounded to the activation dtype
    s_full, z_full = np.repeat(scale, g, axis=1), None if zp is None else np.repeat(zp, g, axis=1)
    W = _round_to(O.fma32(q, s_full, z_full) if zp is not None else (q * s_full).astype(np.float32), dtype)
    ldx = K + 8
    xbuf = np.zeros((M, ldx), dtype=np.uint16)
    xbuf[:, :K] = _bits_of(x, dtype)
    if bias_kind == "none":
        bias, bias_ld, bias_arr = None, 0, None
    elif bias_kind == "vector":
        bias = bias_vals[0].copy()
        bias_ld, bias_arr = 0, np.ascontiguousarray(_bits_of(bias, dtype))
        bias = np.broadcast_to(bias, (M, N))
    else:
        bias, bias_ld, bias_arr = bias_vals, N, np.ascontiguousarray(bias_vals, dtype=np.float32)
    out = np.full((M, N), 0x7FC0 if dtype == "bfloat16" else 0x7E00, dtype=np.uint16)          # NaN-filled: every output must be written
    code = 1 if dtype == "bfloat16" else 2
    bias_code = 0 if bias_kind == "matrix" else code
    for grid in (1, 3):                                                               # one CTA walking all tiles, and a strided grid
        out[:] = 0x7FC0 if dtype == "bfloat16" else 0x7E00
        rc = emu.emu_gemv_packed(_ptr(xbuf), code, ldx, _ptr(raw), ctypes.byref(fmt), _ptr(scale), None if zp is None else _ptr(zp), group,
                                 None if bias_arr is None else _ptr(bias_arr), bias_code, bias_ld, _ptr(out), M, N, K, grid)
        assert rc == 0
        got = _from_bits(out, dtype)
        acc = x.astype(np.float64) @ W.astype(np.float64).T
        if exact:
            want = _round_to((acc.astype(np.float32) + (0 if bias is None else bias.astype(np.float32))).astype(np.float32), dtype)
            np.testing.assert_array_equal(got, want)
        else:
            want = acc + (0 if bias is None else bias)
            # f32 accumulation in the tensor-core model's order + one rounding to T
            bound = np.abs(want) * 2.0 ** (-8 if dtype == "bfloat16" else -11) + (np.abs(x).astype(np.float64) @ np.abs(W).astype(np.float64).T) * 2.0 ** -20 + 1e-6
            assert (np.abs(got - want) <= bound).all(), float(np.abs(got - want).max())


# ---- the Python glue of K5p (forward hook -> ops.linear_small_m_packed -> C ABI arguments), with the C entry point replaced by the
# emulated kernel: layers quantised through the public surface on the CPU, checked against the oracle's dequantise + linear.
SMALL_M_LAYER_CONFIGS = [
    dict(weights_dtype="int4", group_size=128), dict(weights_dtype="uint4"), dict(weights_dtype="int2", group_size=16),
    dict(weights_dtype="float6_e3m2fn", group_size=32), dict(weights_dtype="int5", group_size=-1), dict(weights_dtype="uint7", group_size=8),
    dict(weights_dtype="int4", group_size=-1, use_quantized_matmul=True), dict(weights_dtype="int8", group_size=64, use_quantized_matmul=True),
    dict(weights_dtype="float8_e4m3fn", group_size=32),
]


@pytest.mark.parametrize("cfg", SMALL_M_LAYER_CONFIGS, ids=["-".join(f"{k[:5]}={v}" for k, v in c.items()) for c in SMALL_M_LAYER_CONFIGS])
def test_small_m_packed_forward_glue_on_the_emulator(emu, cfg, monkeypatch):
    import contextlib
    import copy

    import torch

    from sdnq_b200 import SDNQConfig, forward, ops, sdnq_quantize_layer
    emu.emu_gemv_packed.argtypes = [P, ctypes.c_int, ctypes.c_int64, P, ctypes.c_void_p, P, P, ctypes.c_int64, P, ctypes.c_int,
                                    ctypes.c_int64, P, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
    calls = []

    class FakeLib:
        @staticmethod
        def sdnq_b200_linear_small_m_packed(x, x_dtype, ldx, w, fmt, scale, zp, group, bias, bias_dtype, bias_ld, out, M, N, K, stream):
            calls.append((M, N, K, group))
            return emu.emu_gemv_packed(x, x_dtype, ldx, w, ctypes.cast(ctypes.byref(fmt), ctypes.c_void_p), scale, zp, group, bias, bias_dtype,
                                       bias_ld, out, M, N, K, 2)

    monkeypatch.setattr(ops._lib, "load", lambda: FakeLib)
    monkeypatch.setattr(ops, "_require_cuda", lambda *t: None)
    monkeypatch.setattr(ops, "_stream", lambda t: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    monkeypatch.setenv("SDNQ_B200_SMALL_M_PACKED", "1")
    torch.manual_seed(5)
    K, N, M = 256, 72, 5
    lin = torch.nn.Linear(K, N, bias=True).to(torch.bfloat16)
    layer, _ = sdnq_quantize_layer(copy.deepcopy(lin), SDNQConfig(**cfg))
    x = torch.randn(M, K).to(torch.bfloat16)
    assert forward._small_m_packed_ok(layer, x)
    y = forward._small_m_packed_linear(layer, x)
    assert calls and calls[0][:3] == (M, N, K)
    d = layer.sdnq_dequantizer

    def conv(t):
        if t is None:
            return None
        t = t.detach()
        return t.float().numpy() if t.is_floating_point() else t.numpy()
    meta = {k: (list(v) if isinstance(v, (torch.Size, tuple)) else v) for k, v in d.__dict__.items() if k != "result_dtype"}
    ol = O.Layer(conv(layer.weight), conv(layer.scale), conv(layer.zero_point), None, None, bias=conv(layer.bias), **meta)
    want = O.linear_dequant(ol, x.float().numpy(), skip_quantized_matmul=bool(d.use_quantized_matmul))
    got = y.float().numpy()
    assert got.shape == want.shape
    ulp = np.exp2(np.floor(np.log2(np.maximum(np.abs(want), 1e-30))) - 7)
    assert (np.abs(got - want) <= 2 * ulp + 1e-5).all(), float(np.abs(got - want).max())
    # opt-outs: SVD layers, the knob itself
    monkeypatch.setenv("SDNQ_B200_SMALL_M_PACKED", "0")
    assert not forward._small_m_packed_ok(layer, x)


# ---- K2 whole (act_quant_kernel.cuh: argument checks, the K -> (warps per row, chunks per lane) dispatch, tensor-core / butterfly
# rotation, register-resident and two-pass kernels) on emulated CTAs.  launch_pdl is the emulator's grid runner under
# SDNQ_HOST_EMU, so the host dispatch code runs unchanged too.  rcp.approx is modelled by the correctly rounded reciprocal
# (RowDivider refines either to the correctly rounded quotient); everything else is the device source.
def _act_quant_emu(emu, x, dtype, mode, hadamard=0, ldx=None, want_rowsum=False, want_x_rot=False):
    emu.sdnq_b200_act_quant.argtypes = [P, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, P, P, P, P, P, P]
    M, K = x.shape
    ldx = ldx or K
    code = {"float32": 0, "bfloat16": 1, "float16": 2}[dtype]
    if dtype == "float32":
        buf = np.zeros((M, ldx), dtype=np.float32)
        buf[:, :K] = x
    else:
        buf = np.zeros((M, ldx), dtype=np.uint16)
        buf[:, :K] = _bits_of(x, dtype)
    mm = {"int8": 3, "uint8": 4, "fp8": 5}[mode]
    xq = np.full((M, K), 0x55, dtype=np.uint8)
    sx = np.full(M, np.nan, dtype=np.float32)
    zx = np.full(M, np.nan, dtype=np.float32)
    rowsum = np.full(M, -12345, dtype=np.int32)
    x_rot = np.zeros((M, K), dtype=buf.dtype)
    rc = emu.sdnq_b200_act_quant(_ptr(buf), code, M, K, ldx, hadamard, mm, _ptr(xq), _ptr(sx), _ptr(zx) if mode == "uint8" else None,
                                 _ptr(rowsum) if want_rowsum else None, _ptr(x_rot) if want_x_rot else None, None)
    assert rc == 0, rc
    return xq, sx, zx, rowsum, x_rot


def _check_codes(mode, xq, sx, zx, rowsum, x_ref, want_rowsum):
    """x_ref: the (rotated, dtype-rounded) activations the quantiser sees; bit-exact codes / scales / zero points / row sums"""
    if mode == "int8":
        q, s = O.quantize_int_mm(x_ref)
        np.testing.assert_array_equal(xq.view(np.int8), q)
    elif mode == "uint8":
        q, s, z = O.quantize_uint_mm(x_ref)
        np.testing.assert_array_equal(xq.view(np.int8), q)
        np.testing.assert_array_equal(zx, z.reshape(-1))
    else:
        q, s = O.quantize_fp_mm(x_ref)
        np.testing.assert_array_equal(xq, O.e4m3fn_bits(q))
    np.testing.assert_array_equal(sx, s.reshape(-1))
    if want_rowsum:
        np.testing.assert_array_equal(rowsum, xq.view(np.int8).astype(np.int64).sum(axis=1))


@pytest.mark.parametrize("mode", ["int8", "uint8", "fp8"])
@pytest.mark.parametrize("K", [64, 200, 512, 640, 1280, 2560, 5120, 12288, 16640])
def test_act_quant_kernels_on_emulated_ctas(emu, K, mode):
    """every dispatch bucket of K2 (1 / 2 / 4 chunks per lane, 2 / 4 / 8 warps per row, 8 chunks, and the two-pass kernel for
    K > 16384), ragged rows (K % 256 != 0), a padded row stride, an all-zero row: codes, scales, zero points and row sums are
    bit-identical to the reference arithmetic (quant_utils.py:264-299)"""
    rng = np.random.default_rng(K + len(mode))
    M = 3 if K > 4096 else 9                                              # 9 rows: a ragged last CTA for every rows-per-CTA value
    x = O.bf16_round((rng.standard_normal((M, K)) * np.exp2(rng.integers(-6, 6, size=(M, 1)))).astype(np.float32))
    x[1, :] = 0.0                                                         # all-zero row: scale 0, codes 0 (0/0 -> NaN -> 0)
    x[2, rng.integers(0, K)] = 300.0                                      # an outlier channel
    xq, sx, zx, rowsum, _ = _act_quant_emu(emu, x, "bfloat16", mode, ldx=K + 8, want_rowsum=mode != "fp8")
    _check_codes(mode, xq, sx, zx, rowsum, x, want_rowsum=mode != "fp8")


@pytest.mark.parametrize("dtype", ["float32", "float16"])
def test_act_quant_other_activation_dtypes_on_the_emulator(emu, dtype):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5, 1000)).astype(np.float32)
    x = x if dtype == "float32" else x.astype(np.float16).astype(np.float32)
    xq, sx, zx, rowsum, _ = _act_quant_emu(emu, x, dtype, "int8", want_rowsum=True)
    _check_codes("int8", xq, sx, zx, rowsum, x, want_rowsum=True)


@pytest.mark.parametrize("mode", ["int8", "uint8", "fp8"])
@pytest.mark.parametrize("G,K", [(256, 3072), (256, 768), (128, 640), (64, 1344), (16, 528), (32, 17408)])
def test_act_quant_with_rotation_on_emulated_ctas(emu, G, K, mode):
    """K2 with the Hadamard rotation in front (FLUX: fp8 + Hadamard-256; SD-XL C = 640: group 128).  Inputs with exact sums, so the
    rotated activations -- checked through the x_rot output -- and with them codes / scales / row sums must equal the reference's
    rotate_hadamard + quantise bit for bit.  K = 17408 takes the two-pass kernel with the butterfly rotation."""
    rng = np.random.default_rng(G + K)
    M = 2 if K > 4096 else 5
    x = rng.integers(-64, 65, size=(M, K)).astype(np.float32) * np.exp2(rng.integers(-3, 4, size=(M, 1))).astype(np.float32)
    xq, sx, zx, rowsum, x_rot = _act_quant_emu(emu, x, "bfloat16", mode, hadamard=G, want_rowsum=mode != "fp8", want_x_rot=True)
    x_ref = O.rotate_hadamard(x, G, "bfloat16")
    np.testing.assert_array_equal(O.from_bf16_bits(x_rot), x_ref)
    _check_codes(mode, xq, sx, zx, rowsum, x_ref, want_rowsum=mode != "fp8")


def test_small_m_packed_kernel_never_reads_or_writes_past_its_buffers():
    """memcheck on the CPU: every buffer K5p touches ends right before an inaccessible page (tests/host_emu/guarded.py), so an
    overrun of the stored weight, the scales, the activations, the bias or the output kills the (sub)process."""
    import subprocess
    import sys
    import textwrap
    code = textwrap.dedent('''
        import ctypes, sys
        import numpy as np
        sys.path.insert(0, ".")
        from oracle import sdnq_oracle as O
        from tests.host_emu import build_emu
        from tests.host_emu.guarded import guarded
        from tests.test_device_arithmetic_on_host import GEMV_CASES, WeightFormat, _stored_weight, _bits_of, _from_bits, _round_to
        lib = ctypes.CDLL(build_emu.build())
        P = ctypes.c_void_p
        lib.sdnq_b200_linear_small_m_packed.argtypes = [P, ctypes.c_int, ctypes.c_int64, P, ctypes.POINTER(WeightFormat), P, P, ctypes.c_int64, P,
                                                        ctypes.c_int, ctypes.c_int64, P, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, P]
        ran = 0
        for name, group, M, N, K, bias_kind, dtype in GEMV_CASES:
            if (N * K * O.dtype_info(name)["num_bits"] // 8) % 16 or (M * K * 2) % 16:
                continue                                   # the guarded copy must both end at the page boundary and start 16-byte aligned
            rng = np.random.default_rng(ran)
            raw, q, fmt = _stored_weight(rng, name, N, K, False)
            g = group or K
            scale = (rng.random((N, K // g)).astype(np.float32) + 0.5) * np.float32(0.02)
            zp = (rng.standard_normal((N, K // g)).astype(np.float32) * np.float32(0.1)) if O.dtype_info(name)["is_unsigned"] else None
            x = _round_to(rng.standard_normal((M, K)).astype(np.float32), dtype)
            gx, gw, gs = guarded(_bits_of(x, dtype)), guarded(raw), guarded(scale, 4)
            gz = None if zp is None else guarded(zp, 4)
            bias = _round_to(rng.standard_normal(N).astype(np.float32), dtype)
            gb = guarded(_bits_of(bias, dtype), 2) if bias_kind != "none" else None
            gout = guarded(np.zeros((M, N), dtype=np.uint16), 2)
            code = 1 if dtype == "bfloat16" else 2
            rc = lib.sdnq_b200_linear_small_m_packed(gx.ctypes.data, code, K, gw.ctypes.data, ctypes.byref(fmt), gs.ctypes.data,
                                                     None if gz is None else gz.ctypes.data, group, None if gb is None else gb.ctypes.data, code, 0,
                                                     gout.ctypes.data, M, N, K, None)
            assert rc == 0, (name, rc)
            s_full = np.repeat(scale, g, axis=1)
            W = _round_to(O.fma32(q, s_full, np.repeat(zp, g, axis=1)) if zp is not None else (q * s_full).astype(np.float32), dtype)
            want = x.astype(np.float64) @ W.astype(np.float64).T + (0 if gb is None else bias)
            got = _from_bits(np.array(gout), dtype)
            assert np.abs(got - want).max() <= 2e-2 * np.abs(want).max(), name
            ran += 1
        print("guarded cases", ran)
        assert ran >= 8
    ''')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=600)
    assert r.returncode == 0 and "guarded cases" in r.stdout, (r.returncode, r.stdout[-300:], r.stderr[-1500:])
