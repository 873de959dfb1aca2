"""CPU checks of the C-ABI shared object: it loads without a GPU, exports every symbol include/sdnq_b200.h declares, and its
argument validation answers with status codes + messages (no compute is launched here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from sdnq_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sdnq_b200.h")).read()
    return sorted(set(re.findall(r"SDNQ_API\s+[\w\s\*]+?\b(sdnq_b200_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from sdnq_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_workspace_size(lib):
    from sdnq_b200 import _lib
    assert lib.sdnq_b200_abi_version() == _lib.ABI_VERSION
    assert lib.sdnq_b200_linear_w8a8_workspace_bytes(0, 640) == 0
    need = lib.sdnq_b200_linear_w8a8_workspace_bytes(4096, 640)
    assert need >= 4096 * 640 + 3 * 4096 * 4 and need % 256 == 0


def test_argument_errors_are_reported_not_thrown(lib):
    from sdnq_b200._lib import WeightFormat
    fmt = WeightFormat(0, 12, 0, 0, 0, 1)                      # 12-bit: valid upstream dtype, no CUDA kernel
    rc = lib.sdnq_b200_unpack(ctypes.c_void_p(16), ctypes.byref(fmt), ctypes.c_void_p(16), 3, 8, None)
    assert rc == -2 and b"8 bits" in lib.sdnq_b200_last_error()
    fmt = WeightFormat(1, 6, 0, 3, 3, 1)                       # sign + 3 + 3 != 6
    rc = lib.sdnq_b200_unpack(ctypes.c_void_p(16), ctypes.byref(fmt), ctypes.c_void_p(16), 0, 8, None)
    assert rc == -1 and b"minifloat" in lib.sdnq_b200_last_error()
    rc = lib.sdnq_b200_scaled_mm(ctypes.c_void_p(16), ctypes.c_void_p(16), 3, None, None, None, 0, 0, None, None, None, None,
                                 ctypes.c_void_p(16), 1, 64, 64, 24, None)
    assert rc == -2 and b"multiple of 16" in lib.sdnq_b200_last_error()
    rc = lib.sdnq_b200_act_quant(None, 1, 4, 64, 64, 0, 3, None, None, None, None, None, None)
    assert rc == -1


def test_argument_errors_of_the_conv_and_small_m_entries(lib):
    """the entry points added for the conv path and the small-M Linear validate before touching CUDA"""
    from sdnq_b200._lib import Conv2dGeometry, WeightFormat
    P = ctypes.c_void_p
    rc = lib.sdnq_b200_linear_small_m(P(16), 1, 64, P(16), 3, P(16), None, None, 0, P(16), 40, 64, 64, None)          # M > 32
    assert rc == -1 and b"M <= 32" in lib.sdnq_b200_last_error()
    rc = lib.sdnq_b200_linear_small_m(P(16), 0, 64, P(16), 3, P(16), None, None, 0, P(16), 4, 64, 64, None)           # f32 activations
    assert rc == -2 and b"bf16 / f16" in lib.sdnq_b200_last_error()
    rc = lib.sdnq_b200_linear_small_m(P(16), 1, 72, P(16), 3, P(16), None, None, 0, P(16), 4, 64, 72, None)           # K % 16
    assert rc == -2
    rc = lib.sdnq_b200_rows_to_nchw(P(16), P(16), 3, 1, 64, 64, None)                                                  # 3-byte elements
    assert rc == -1 and b"2 or 4 bytes" in lib.sdnq_b200_last_error()
    geo = Conv2dGeometry(1, 64, 8, 8, 4096, 64, 8, 1, 3, 3, 0, 1, 1, 1, 1, 1)                                           # stride_h = 0
    rc = lib.sdnq_b200_conv_act_quant(P(16), 1, ctypes.byref(geo), 0, 3, P(16), P(16), None, None, None, None)
    assert rc == -1 and b"geometry" in lib.sdnq_b200_last_error()
    geo = Conv2dGeometry(2, 64, 8, 8, 4096, 64, 8, 1, 3, 3, 1, 1, 1, 1, 1, 1)
    assert lib.sdnq_b200_conv_act_quant_workspace_bytes(ctypes.byref(geo), 3) == 2 * 8 * 8 * 4
    assert lib.sdnq_b200_conv_act_quant_workspace_bytes(ctypes.byref(geo), 4) == 2 * 8 * 8 * 4 * 2                        # uint8: min and max
    rc = lib.sdnq_b200_conv_act_quant_ws(P(16), 1, ctypes.byref(geo), 0, 3, P(16), P(16), None, None, None, P(16), 8, None)    # workspace too small
    assert rc == -1 and b"workspace" in lib.sdnq_b200_last_error()
    fmt = WeightFormat(0, 8, 0, 0, 0, 1)
    dims = (ctypes.c_int64 * 2)(3, 5)                                                                                    # 15 weights: not a multiple of 8
    strides = (ctypes.c_int64 * 2)(1, 0)
    rc = lib.sdnq_b200_dequant_nd(P(16), ctypes.byref(fmt), P(16), None, 0, 2, dims, strides, None, 0, P(16), 1, None)
    assert rc == -2 and b"multiple of 8" in lib.sdnq_b200_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from sdnq_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.SDNQKernelError, match="no CPU or eager fallback"):
        _lib.load()
