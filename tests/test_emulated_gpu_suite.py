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
        values = [v.values if hasattr(v, "values") else v for v in mark.args[1]]
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
    return "-".join(os.path.basename(v)[6:-4] if isinstance(v, str) and v.endswith(".npz") else str(v).replace("torch.", "") for v in kw.values())


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
