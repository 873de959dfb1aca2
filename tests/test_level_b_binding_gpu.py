"""Level-B drop-in proof on the GPU: the UNMODIFIED reference keeps its own module tree, quantiser and
`SDNQLinear.forward`; only its kernel layer is re-pointed at this library through the exact ctypes stub printed in INTEGRATION.md
(`sdnq/kernels/b200.py`: `sdnq_scaled_mm` over `sdnq_b200_scaled_mm`, `quantize_int_mm_input` over `sdnq_b200_act_quant`).  The
outputs of the reference's forward with the stub installed are compared with the reference's own CUDA-eager output
(torch._int_mm / torch._scaled_mm, kernel_wrappers.py:132-150) on the same layer and input.

Runs in a child process: the reference resolves its flags at import time (SDNQ_DEVICE, SDNQ_USE_TRITON_MM ...).  Needs
oracle/_ref (a scripted copy of the reference package, `python oracle/build_ref.py`; it travels to the GPU box with the snapshot)."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = textwrap.dedent('''
    import copy, json, os, re, sys
    sys.path.insert(0, os.getcwd())
    import torch
    from oracle.ref_loader import load_reference
    sdnq = load_reference(SDNQ_DEVICE="cuda", SDNQ_USE_TORCH_COMPILE="0", SDNQ_USE_TRITON_MM="0")
    from sdnq import SDNQConfig, kernel_wrappers
    from sdnq.quantizer import sdnq_quantize_layer
    from sdnq.layers.linear import linear_int8
    from sdnq_b200 import _lib
    assert kernel_wrappers.sdnq_scaled_mm is None, "the oracle side must be the reference's CUDA-eager path"
    md = open("INTEGRATION.md").read()
    stub_src = next(b for b in re.findall(r"```python\\n(.*?)```", md, flags=re.S) if "sdnq/kernels/b200.py" in b)
    stub = {}
    exec(compile(stub_src.replace('"libsdnq_b200.so"', repr(_lib.LIB_PATH)), "INTEGRATION.md:b200.py", "exec"), stub)

    def ulp(a, b):
        def key(t):
            i = t.contiguous().view(torch.int16).to(torch.int32) & 0xFFFF
            return torch.where((i & 0x8000) != 0, -(i & 0x7FFF), i)
        return (key(a) - key(b)).abs()

    out = {}
    cases = {"int8": (dict(weights_dtype="int8", use_quantized_matmul=True), 256, 640, 1280, True),
             "int8_nobias": (dict(weights_dtype="int8", use_quantized_matmul=True), 1024, 1280, 1280, False),
             "fp8": (dict(weights_dtype="float8_e4m3fn", use_quantized_matmul=True), 256, 1024, 768, True)}
    for name, (cfg, M, K, N, bias) in cases.items():
        torch.manual_seed(len(name) + M)
        lin = torch.nn.Linear(K, N, bias=bias).to(device="cuda", dtype=torch.bfloat16)
        layer = sdnq_quantize_layer(copy.deepcopy(lin), SDNQConfig(**cfg))[0]
        x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        with torch.no_grad():
            y_ref = layer(x)                                  # reference kernels (CUDA eager)
            saved = (kernel_wrappers.sdnq_scaled_mm, linear_int8.quantize_int_mm_input)
            kernel_wrappers.sdnq_scaled_mm = stub["sdnq_scaled_mm"]                  # kernel_wrappers.py:193-204 now reach libsdnq_b200.so
            linear_int8.quantize_int_mm_input = stub["quantize_int_mm_input"]        # linear_int8.py:14-22
            n0 = _lib.launch_count(reset=True)
            y_ours = layer(x)
            launches = _lib.launch_count()
            kernel_wrappers.sdnq_scaled_mm, linear_int8.quantize_int_mm_input = saved
        du = ulp(y_ours, y_ref)
        out[name] = {"forward": layer.forward_func.__name__, "launches": int(launches), "max_ulp": int(du.max()), "frac_diff": float((du > 0).float().mean()),
                     "shape_ok": y_ours.shape == y_ref.shape and y_ours.dtype == y_ref.dtype, "ref_module": type(layer).__module__}
    print("LEVEL_B " + json.dumps(out))
''')


def test_reference_forward_over_the_c_abi_stub_matches_reference_cuda_eager():
    sys.path.insert(0, ROOT)
    from oracle.ref_loader import reference_root
    if reference_root() is None:
        pytest.skip("oracle/_ref is not present (python oracle/build_ref.py in the authoring container)")
    env = dict(os.environ)
    for k in ("SDNQ_USE_TRITON_MM", "SDNQ_USE_TORCH_COMPILE", "SDNQ_DEVICE"):
        env.pop(k, None)
    p = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    got = [ln for ln in p.stdout.splitlines() if ln.startswith("LEVEL_B ")]
    assert p.returncode == 0 and got, (p.stdout[-1500:], p.stderr[-3000:])
    res = json.loads(got[-1][len("LEVEL_B "):])
    for name, r in res.items():
        assert r["ref_module"].startswith("sdnq."), "the module tree must be the reference's own"
        assert r["shape_ok"], name
        # int8: K2 + K1 through the stub; fp8: the reference's own activation quantiser + K1
        assert r["launches"] == (2 if name.startswith("int8") else 1), (name, r)
        # exact integer / fp8 contraction on both sides, same activation codes: only the f32 epilogue order can move a bf16 ulp
        assert r["max_ulp"] <= 1 and r["frac_diff"] < 0.02, (name, r)
