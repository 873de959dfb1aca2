"""The GPU kernel-parity tests, re-run on the CPU through the host emulation of libsdnq_b200.so.

`tests/host_emu/build_emu.py` compiles the CUDA sources of every kernel family that does not need tcgen05 / TMA -- K2 (act_quant*.cu),
K2c (conv im2col quantisers), K3 / K4 / unpack (dequant.cu), K3c (dequant_nd.cu), K5 (gemv_w8a16.cu), K5p (gemv_packed.cu) -- unchanged
with g++; `launch_pdl` / `launch_plain` run the kernels on lock-stepped host threads (warp.h) and the result exports the same C ABI.
This file points `sdnq_b200._lib.load()` at that library, lets `sdnq_b200.ops` accept CPU tensors, and calls the test functions of
`tests/test_kernels_gpu.py` / `tests/test_conv_gpu.py` themselves (same fixtures, same assertions, same parameter lists, thinned
where the emulator would take minutes).  The tcgen05 GEMM (K1) and the tensor-core SVD update (K3s) stay GPU-only.

What this proves: host dispatch code, indexing, fragment mappings, reductions and arithmetic of those kernels against the
reference-generated fixtures and the oracle, on every driver CPU run.  What it cannot prove: hardware behaviour of the three
approximated primitives (rcp.approx is modelled by the exact reciprocal), alignment faults, races that lock-stepping hides,
performance.  The `-m gpu` tests remain the parity tests proper."""
import contextlib
import ctypes
import itertools

import pytest
import torch

from tests import test_conv_gpu as C
from tests import test_kernels_gpu as K
from tests.host_emu import build_emu


@pytest.fixture(scope="module", autouse=True)
def emulated_library():
    from sdnq_b200 import _lib, ops
    lib = ctypes.CDLL(build_emu.build())
    for name, (res, args) in _lib.SIGNATURES.items():
        if name in build_emu.NOT_EMULATED:
            continue
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    mp = pytest.MonkeyPatch()
    mp.setattr(_lib, "load", lambda: lib)
    mp.setattr(ops, "_require_cuda", lambda *t: None)
    mp.setattr(ops, "_stream", lambda t: 0)
    mp.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    mp.setattr(K, "DEV", "cpu")
    mp.setattr(C, "DEV", "cpu")
    yield lib
    mp.undo()


def cases(func, keep=None):
    """the cartesian product of a test function's own @pytest.mark.parametrize lists, as kwargs dicts (optionally filtered)"""
    axes = []
    for mark in getattr(func, "pytestmark", []):
        if mark.name != "parametrize":
            continue
        names = [n.strip() for n in mark.args[0].split(",")]
        values = [v.values if hasattr(v, "marks") else v for v in mark.args[1]]      # pytest.param(...) -> its values
        axes.append([dict(zip(names, v if len(names) > 1 else (v,))) for v in values])
    out = []
    for combo in itertools.product(*axes):
        kw = {}
        for part in combo:
            kw.update(part)
        if keep is None or keep(kw):
            out.append(kw)
    return out


def ident(kw):
    import os
    def one(v):
        if isinstance(v, str) and v.endswith(".npz"):
            return os.path.basename(v)[6:-4]
        if isinstance(v, dict):
            return "_".join(str(x) for x in v.values())
        return str(v).replace("torch.", "")
    return "-".join(one(v) for v in kw.values())


def emulated(func, keep=None):
    params = cases(func, keep)
    assert params, func.__name__
    return pytest.mark.parametrize("kw", params, ids=[ident(p) for p in params])


def test_every_other_entry_point_is_emulated(emulated_library):
    from sdnq_b200 import _lib
    for name in _lib.SIGNATURES:
        assert hasattr(emulated_library, name) != (name in build_emu.NOT_EMULATED), name


# ---- unpack (dequant.cu: unpack_kernel)
@emulated(K.test_unpack_int_bit_exact)
def test_unpack_int(kw):
    if kw["bits"] == 1 and kw["signed"]:
        pytest.skip("int1 is an alias of uint1")
    K.test_unpack_int_bit_exact(**kw)


def test_unpack_misc():
    K.test_unpack_uint1_int64_words()
    K.test_unpack_golden_kat()
    K.test_unpack_minifloat_all_codes()


# ---- K3 / K4 on the reference-generated layer fixtures
@emulated(K.test_dequant_matches_reference)
def test_dequant_fixtures(kw):
    K.test_dequant_matches_reference(**kw)


@emulated(K.test_requant_matches_reference_bit_exact)
def test_requant_fixtures(kw):
    K.test_requant_matches_reference_bit_exact(**kw)


@emulated(K.test_dequant_flat_path_bit_exact, keep=lambda kw: kw["N"] * kw["K"] <= 600000)
def test_dequant_flat_path(kw):
    K.test_dequant_flat_path_bit_exact(**kw)


@emulated(K.test_dequant_rotated_8bit, keep=lambda kw: kw["N"] <= 96)
def test_dequant_rotated_8bit(kw):
    K.test_dequant_rotated_8bit(**kw)


# ---- K2
@emulated(K.test_act_quant_bit_exact)
def test_act_quant(kw):
    K.test_act_quant_bit_exact(**kw)


@emulated(K.test_act_quant_other_activation_dtypes)
def test_act_quant_dtypes(kw):
    K.test_act_quant_other_activation_dtypes(**kw)


@emulated(K.test_act_quant_long_rows)
def test_act_quant_long_rows(kw):
    K.test_act_quant_long_rows(**kw)


@emulated(K.test_hadamard_rotation_matches_oracle)
def test_hadamard_rotation(kw):
    K.test_hadamard_rotation_matches_oracle(**kw)


@emulated(K.test_hadamard_f16_and_partial_chunks)
def test_hadamard_f16_partial(kw):
    K.test_hadamard_f16_and_partial_chunks(**kw)


# ---- K5
@emulated(K.test_small_m_linear, keep=lambda kw: kw["M"] in (1, 13, 32) and kw["K"] <= 1280)
def test_small_m_linear(kw):
    K.test_small_m_linear(**kw)


def test_small_m_linear_strided_and_errors():
    K.test_small_m_linear_strided_and_errors()


# ---- convolutions: K3c weight dequant, K2c im2col quantisers
@emulated(C.test_conv_weight_dequant_matches_reference)
def test_conv_weight_dequant(kw):
    C.test_conv_weight_dequant_matches_reference(**kw)


@emulated(C.test_conv_act_quant_matches_reference)
def test_conv_act_quant(kw):
    C.test_conv_act_quant_matches_reference(**kw)


# ---- layer level (SDNQConfig -> sdnq_quantize_layer -> forward_func -> ops -> C ABI) for every forward that does not need the tcgen05
# GEMM: the dequant path (K3 + the library matmul, here torch's CPU matmul), the rows < 32 branches (K5, K5p, dequantise + matmul)
from tests import test_layers_gpu as L  # noqa: E402
from tests import test_small_m_packed_gpu as Z  # noqa: E402
from tests.util import fixture_tensors  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def layer_modules_on_cpu(emulated_library):
    mp = pytest.MonkeyPatch()
    mp.setattr(L, "DEV", "cpu")
    mp.setattr(Z, "DEV", "cpu")
    yield
    mp.undo()


@emulated(L.test_forward_matches_reference_output)
def test_layer_forward_fixtures(kw):
    _, _, meta = fixture_tensors(kw["path"])
    if meta["dequantizer"]["use_quantized_matmul"] and meta["M"] >= 32:
        pytest.skip("W8A8 with 32 or more rows runs the tcgen05 GEMM: GPU only")
    L.test_forward_matches_reference_output(**kw)


def _thin_small_m(kw):
    """rotated layers run the rotated K3 over the whole weight as their reference path (seconds on the emulator): one M for those"""
    return kw["M"] == 4 if (kw["cfg"].get("use_hadamard") or kw["cfg"].get("use_svd")) else kw["M"] in (1, 31)


@emulated(L.test_small_m_forward_gemv_vs_dequant_path, keep=_thin_small_m)
def test_small_m_forward_gemv(kw, monkeypatch):
    L.test_small_m_forward_gemv_vs_dequant_path(monkeypatch=monkeypatch, **kw)


@emulated(Z.test_small_m_packed_forward_vs_dequant_path, keep=_thin_small_m)
def test_small_m_packed_forward(kw, monkeypatch):
    """K5p at layer level: the very test that is its hardware gate (tests/test_small_m_packed_gpu.py), on the emulator"""
    Z.test_small_m_packed_forward_vs_dequant_path(monkeypatch=monkeypatch, **kw)


@emulated(Z.test_small_m_packed_forward_vs_oracle)
def test_small_m_packed_forward_vs_oracle(kw):
    Z.test_small_m_packed_forward_vs_oracle(**kw)


def test_emulator_traps_misaligned_vector_access():
    """x86 would silently execute the misaligned 16-byte loads a GPU faults on; the emulator is built with -fsanitize=alignment, so a
    kernel that issues one aborts the process.  Shown here by handing K5p's body (which skips the entry point's own alignment check)
    an activation pointer that is 2 bytes off."""
    import subprocess
    import sys
    import textwrap
    code = textwrap.dedent('''
        import ctypes, sys
        import numpy as np
        sys.path.insert(0, ".")
        from tests.host_emu import build_emu
        lib = ctypes.CDLL(build_emu.build())
        P = ctypes.c_void_p
        class WF(ctypes.Structure):
            _fields_ = [(n, ctypes.c_int32) for n in ("kind", "bits", "is_unsigned", "exponent", "mantissa", "word_bytes")]
        M, N, K = 2, 16, 64
        x = np.zeros(M * K + 8, dtype=np.uint16); w = np.zeros(N * K // 2 + 16, dtype=np.uint8)
        s = np.ones(N, dtype=np.float32); out = np.zeros(M * N, dtype=np.uint16)
        lib.emu_gemv_packed.argtypes = [P, ctypes.c_int, ctypes.c_int64, P, P, P, P, ctypes.c_int64, P, ctypes.c_int, ctypes.c_int64, P,
                                        ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        fmt = WF(0, 4, 0, 0, 0, 1)
        rc = lib.emu_gemv_packed(x.ctypes.data + int(sys.argv[1]), 1, K, w.ctypes.data, ctypes.addressof(fmt), s.ctypes.data, None, 0, None, 0, 0,
                                 out.ctypes.data, M, N, K, 1)
        print("rc", rc)
    ''')
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ok = subprocess.run([sys.executable, "-c", code, "0"], capture_output=True, text=True, cwd=root, timeout=300)
    assert ok.returncode == 0 and "rc 0" in ok.stdout, ok.stderr[-500:]
    bad = subprocess.run([sys.executable, "-c", code, "2"], capture_output=True, text=True, cwd=root, timeout=300)
    assert bad.returncode != 0 and "misaligned" in bad.stderr, (bad.returncode, bad.stderr[-500:])


@emulated(C.test_conv_forward_matches_reference_output)
def test_conv_forward_fixtures(kw):
    """quantized conv forwards at layer level: the dequant path (K3c + the library convolution, here torch's CPU convolution) and
    the rows < 32 branch; W8A8 convolutions reach the tcgen05 GEMM, which the emulated library does not export"""
    try:
        C.test_conv_forward_matches_reference_output(**kw)
    except AttributeError as e:
        if "scaled_mm" not in str(e):
            raise
        pytest.skip("W8A8 convolution: runs the tcgen05 GEMM, GPU only")


@emulated(L.test_dequantize_restores_module)
def test_dequantize_restores_module(kw):
    try:
        L.test_dequantize_restores_module(**kw)
    except AttributeError as e:
        if "scaled_mm" not in str(e) and "linear_w8a8" not in str(e):
            raise
        pytest.skip("the quantised forward of this fixture runs the tcgen05 GEMM: GPU only")


@emulated(L.test_dynamic_quantization_picks_the_reference_dtypes)
def test_dynamic_quantization(kw):
    """use_dynamic_quantization: the per-layer dtype search dequantises every trial through K3 -- the dtypes it picks equal the
    reference's run (tests/golden/model_dynamic.json) on the emulator as on the GPU"""
    try:
        L.test_dynamic_quantization_picks_the_reference_dtypes(**kw)
    except AttributeError as e:
        if "scaled_mm" not in str(e) and "linear_w8a8" not in str(e):
            raise
        pytest.skip("the model forward at the end of this case runs the tcgen05 GEMM: GPU only")


# ---- K7 (svd_low.cu): the rank-r projection of the W8A8 SVD branch
@emulated(K.test_svd_low, keep=lambda kw: kw["M"] * kw["K"] <= 130 * 2048)
def test_svd_low(kw):
    K.test_svd_low(**kw)



# ---- K8 (weight_quant.cu): load-time quantise + pack against the host arithmetic and the reference fixtures
@emulated(K.test_quantize_weight_matches_host_arithmetic,
          keep=lambda kw: kw["N"] * kw["K"] <= 33 * 640 and (kw["K"] != 1536 or kw["wd"] in ("int4", "uint8"))
          and (kw["gs"] > 0 or kw["wd"] in ("int8", "uint8", "int4", "uint4", "int5", "uint3")))      # (the CTA-per-group kernel takes seconds per case here)
def test_quantize_weight(kw):
    K.test_quantize_weight_matches_host_arithmetic(**kw)


_EMU_FLOATS = ("float8_e4m3fn", "float8_e5m2", "float8_e4m3fn_sdnq", "float7_e2m5fnu", "float6_e3m2fn", "float5_e4m0fn", "float4_e2m1fn", "float4_e2m2fnu",
               "float3_e1m1fn", "float2_e1m0fn")


@emulated(K.test_quantize_weight_float_formats_match_host_arithmetic,      # a representative third of the formats; the GPU run takes all twenty
          keep=lambda kw: kw["wd"] in _EMU_FLOATS and kw["N"] * kw["K"] <= 33 * 640
          and (kw["K"] != 1536 or kw["wd"] in ("float6_e3m2fn", "float8_e4m3fn")) and (kw["gs"] > 0 or kw["wd"] in ("float8_e4m3fn", "float6_e3m2fn", "float4_e2m2fnu")))
def test_quantize_weight_float_formats(kw):
    K.test_quantize_weight_float_formats_match_host_arithmetic(**kw)


@emulated(K.test_quantize_weight_reproduces_reference_fixture)
def test_quantize_weight_fixture(kw):
    K.test_quantize_weight_reproduces_reference_fixture(**kw)


# ---- quantized embedding lookup (K3 with a row gather)
@emulated(L.test_quantized_embedding_forward)
def test_quantized_embedding(kw):
    if kw["cfg"].get("use_svd"):
        pytest.skip("the full-table reference of SVD layers goes through the tcgen05 SVD kernel on the GPU; on the emulator it takes minutes")
    L.test_quantized_embedding_forward(**kw)
