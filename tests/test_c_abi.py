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
    fmt4 = WeightFormat(0, 4, 0, 0, 0, 1)
    args = lambda M, K, group, fmt=fmt4, ld=0: (P(16), 1, K, P(16), ctypes.byref(fmt), P(16), None, group, None, 0, ld, P(16), M, 64, K, None)  # noqa: E731
    rc = lib.sdnq_b200_linear_small_m_packed(*args(40, 64, 0))                                                         # M > 32
    assert rc == -1 and b"M <= 32" in lib.sdnq_b200_last_error()
    rc = lib.sdnq_b200_linear_small_m_packed(*args(4, 64, 12))                                                         # group not a multiple of 8
    assert rc == -2 and b"multiple of 8" in lib.sdnq_b200_last_error()
    rc = lib.sdnq_b200_linear_small_m_packed(*args(4, 64, 0, fmt=WeightFormat(0, 1, 1, 0, 0, 8)))                      # uint1 in int64 words
    assert rc == -2 and b"int64" in lib.sdnq_b200_last_error()
    rc = lib.sdnq_b200_linear_small_m_packed(*args(4, 64, 0, ld=8))                                                    # matrix bias narrower than N
    assert rc == -1 and b"bias_ld" in lib.sdnq_b200_last_error()
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


# ---- the header is the contract: the ctypes binding (and the stub INTEGRATION.md shows to reference maintainers) must agree with it
def header_prototypes():
    """{name: (return type, [parameter types])} parsed from include/sdnq_b200.h (comments stripped)."""
    text = open(os.path.join(ROOT, "include", "sdnq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for ret, name, params in re.findall(r"SDNQ_API\s+([\w\s\*]+?)\b(sdnq_b200_\w+)\s*\(([^)]*)\)\s*;", text):
        plist = []
        for p in params.split(","):
            p = " ".join(p.split())
            if p in ("void", ""):
                continue
            m = re.match(r"(.*?)(\w+)$", p)                    # drop the parameter name
            plist.append(m.group(1).replace(" ", ""))
        protos[name] = (ret.replace(" ", ""), plist)
    return protos, text


def _ctype_of(c_type):
    from sdnq_b200._lib import Conv2dGeometry, WeightFormat
    table = {"int": ctypes.c_int, "float": ctypes.c_float, "int64_t": ctypes.c_int64, "size_t": ctypes.c_size_t, "constchar*": ctypes.c_char_p,
             "constsdnq_weight_format*": ctypes.POINTER(WeightFormat), "constsdnq_conv2d_geometry*": ctypes.POINTER(Conv2dGeometry),
             "constint64_t*": ctypes.POINTER(ctypes.c_int64), "void*const*": ctypes.POINTER(ctypes.c_void_p)}
    if c_type in table:
        return table[c_type]
    assert c_type.endswith("*"), f"unmapped C type {c_type!r}"
    return ctypes.c_void_p                                     # every other pointer (void / float / int32_t, const or not) is a device pointer


def test_ctypes_signatures_match_the_header_types():
    from sdnq_b200 import _lib
    protos, _ = header_prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    for name, (ret, params) in protos.items():
        res, args = _lib.SIGNATURES[name]
        assert res is _ctype_of(ret), f"{name}: return type {ret} bound as {res}"
        assert len(args) == len(params), f"{name}: {len(params)} parameters in the header, {len(args)} in the binding"
        for i, (c, a) in enumerate(zip(params, args)):
            assert a is _ctype_of(c), f"{name}: parameter {i} is {c} in the header, {a} in the binding"


def test_struct_layouts_match_the_header():
    from sdnq_b200._lib import Conv2dGeometry, DequantJob, WeightFormat
    _, text = header_prototypes()
    ctype = {"int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "int": ctypes.c_int, "sdnq_weight_format": WeightFormat}
    for struct, cls in (("sdnq_weight_format", WeightFormat), ("sdnq_conv2d_geometry", Conv2dGeometry), ("sdnq_dequant_job", DequantJob)):
        body = re.search(r"typedef\s+struct\s+" + struct + r"\s*\{(.*?)\}\s*" + struct + r"\s*;", text, flags=re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if decl:
                m = re.match(r"(.*?[\s\*])(\w+(?:\s*,\s*\w+)*)$", decl)          # type, then one or more names
                typ = m.group(1).replace(" ", "")
                fields += [(n.strip(), ctypes.c_void_p if typ.endswith("*") else ctype[typ]) for n in m.group(2).split(",")]
        assert [(n, t) for n, t in cls._fields_] == fields, f"{struct}: ctypes fields differ from the header"


def test_integration_stub_binds_against_the_built_library(lib):
    """The ctypes stub INTEGRATION.md proposes for the reference (sdnq/kernels/b200.py) loads our library, and the argument
    lists it declares have the header's length and types."""
    from sdnq_b200 import _lib
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", md, flags=re.S)
    stub = next(b for b in blocks if "sdnq/kernels/b200.py" in b)
    assert '"libsdnq_b200.so"' in stub
    ns = {}
    exec(compile(stub.replace('"libsdnq_b200.so"', repr(_lib.LIB_PATH)), "INTEGRATION.md:b200.py", "exec"), ns)
    exec(compile(next(b for b in blocks if "sdnq/kernels/b200_atten.py" in b), "INTEGRATION.md:b200_atten.py", "exec"), ns)      # the attention launch
    protos, _ = header_prototypes()
    bound = 0
    for name, (_, params) in protos.items():
        fn = getattr(ns["_lib"], name)
        if fn.argtypes is None:
            continue
        bound += 1
        assert len(fn.argtypes) == len(params), f"INTEGRATION.md stub: {name} declares {len(fn.argtypes)} arguments, header has {len(params)}"
        for i, (c, a) in enumerate(zip(params, fn.argtypes)):
            want = _ctype_of(c)
            want = ctypes.c_void_p if isinstance(want, type) and issubclass(want, ctypes._Pointer) else want
            assert a is want, f"INTEGRATION.md stub: {name} parameter {i}: {c} bound as {a}"
    assert bound >= 4 and ns["_lib"].sdnq_b200_attention.argtypes is not None
    for fn in ("sdnq_scaled_mm", "quantize_int_mm_input", "sdnq_atten_fwd"):
        assert callable(ns[fn])
    assert [ns[k] for k in ("F32", "BF16", "F16", "I8", "U8", "F8E4M3", "I32")] == \
        [_lib.SDNQ_F32, _lib.SDNQ_BF16, _lib.SDNQ_F16, _lib.SDNQ_I8, _lib.SDNQ_U8, _lib.SDNQ_F8E4M3, _lib.SDNQ_I32]


def test_enum_values_match_the_header():
    from sdnq_b200 import _lib
    _, text = header_prototypes()
    values = {}
    for body in re.findall(r"enum\s*\w*\s*\{(.*?)\}", text, flags=re.S):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            name, _, val = (s.strip() for s in item.partition("="))
            nxt = int(val, 0) if val else nxt
            values[name] = nxt
            nxt += 1
    for name in ("SDNQ_F32", "SDNQ_BF16", "SDNQ_F16", "SDNQ_I8", "SDNQ_U8", "SDNQ_F8E4M3", "SDNQ_I32",
                 "SDNQ_W_INT", "SDNQ_W_MINIFLOAT", "SDNQ_W_FP8_E4M3FN", "SDNQ_W_FP8_E5M2"):
        assert values[name] == getattr(_lib, name), name


def test_plain_c_client_links_and_runs(lib, tmp_path):
    """include/sdnq_b200.h is valid C99 (-pedantic -Werror), and a program with no torch / CUDA headers links against the
    library and gets status codes + messages back (tests/c_abi/c_client.c; every call returns before touching CUDA)."""
    import subprocess

    from sdnq_b200 import _lib
    exe = str(tmp_path / "c_client")
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c_abi", "c_client.c"), "-o", exe, "-L", libdir, "-lsdnq_b200", f"-Wl,-rpath,{libdir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "0 failure(s)" in r.stdout, r.stdout + r.stderr


def test_last_error_is_per_thread(lib):
    """exported functions must be callable from several Python threads; the error string is thread-local,
    so concurrent failing calls never see each other's message."""
    import threading

    from sdnq_b200._lib import WeightFormat
    wide = WeightFormat(0, 12, 0, 0, 0, 1)
    start = threading.Barrier(4)
    bad = []

    def worker(kind):
        start.wait()
        for _ in range(2000):
            if kind % 2:
                rc = lib.sdnq_b200_unpack(ctypes.c_void_p(16), ctypes.byref(wide), ctypes.c_void_p(16), 3, 8, None)
                want = b"8 bits"
            else:
                rc = lib.sdnq_b200_rows_to_nchw(ctypes.c_void_p(16), ctypes.c_void_p(16), 3, 1, 64, 64, None)
                want = b"2 or 4 bytes"
            msg = lib.sdnq_b200_last_error()
            if rc >= 0 or want not in msg:
                bad.append((kind, rc, msg))
                return

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not bad, bad[:3]
