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
            # (1) only the contraction re-pointed: kernel_wrappers.py:193-204 now reach libsdnq_b200.so, the activation codes are
            #     still the reference's own
            kernel_wrappers.sdnq_scaled_mm = stub["sdnq_scaled_mm"]
            _lib.launch_count(reset=True)
            y_k1 = layer(x)
            launches_k1 = _lib.launch_count()
            # (2) the activation quantiser re-pointed as well (linear_int8.py:14-22)
            linear_int8.quantize_int_mm_input = stub["quantize_int_mm_input"]
            _lib.launch_count(reset=True)
            y_ours = layer(x)
            launches = _lib.launch_count()
            kernel_wrappers.sdnq_scaled_mm, linear_int8.quantize_int_mm_input = saved
            r = {}
            if name.startswith("int8"):
                xq_r, sx_r = linear_int8.quantize_int_mm_input(x, dtype=torch.float32)
                xq_o, sx_o = stub["quantize_int_mm_input"](x, dtype=torch.float32)
                amax = x.float().abs().amax(dim=-1, keepdim=True)
                same_row = (sx_r == sx_o).flatten() & (xq_r == xq_o).all(dim=-1)
                r.update(rows=int(M), rows_same_codes_and_scale=int(same_row.sum()), scales_differ=int((sx_r != sx_o).sum()),
                         codes_differ=int((xq_r != xq_o).sum()), max_code_delta=int((xq_r.int() - xq_o.int()).abs().max()),
                         ref_scale_is_amax_times_rcp127=bool(torch.equal(sx_r, amax * (1.0 / 127))),
                         ref_scale_is_amax_div_127=bool(torch.equal(sx_r, amax / torch.full_like(amax, 127.0))),
                         our_scale_is_amax_div_127=bool(torch.equal(sx_o, amax / torch.full_like(amax, 127.0))),
                         max_ulp_rows_same=int(ulp(y_ours, y_ref)[same_row].max()) if bool(same_row.any()) else 0)
        du1, du = ulp(y_k1, y_ref), ulp(y_ours, y_ref)
        scale = float(y_ref.float().abs().max())
        r.update(forward=layer.forward_func.__name__, launches_k1=int(launches_k1), launches=int(launches),
                 k1_max_ulp=int(du1.max()), k1_frac_diff=float((du1 > 0).float().mean()),
                 max_ulp=int(du.max()), frac_diff=float((du > 0).float().mean()),
                 max_abs_err_over_range=float((y_ours.float() - y_ref.float()).abs().max()) / scale,
                 shape_ok=bool(y_ours.shape == y_ref.shape and y_ours.dtype == y_ref.dtype), ref_module=type(layer).__module__)
        out[name] = r
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
    print(json.dumps(res, indent=1))
    for name, r in res.items():
        assert r["ref_module"].startswith("sdnq."), "the module tree must be the reference's own"
        assert r["shape_ok"], name
        # (1) the contraction alone over the C ABI: exact integer / fp8 products on both sides and the same activation codes, so
        #     only the f32 epilogue order can move a bf16 ulp
        assert r["launches_k1"] == 1, (name, r)
        assert r["k1_max_ulp"] <= 1 and r["k1_frac_diff"] < 0.02, (name, r)
        # (2) with the activation quantiser re-pointed too.  This library divides (amax / 127, true IEEE division: the reference's
        #     CPU arithmetic, which the committed fixtures pin); the reference's CUDA-eager path multiplies by the rounded reciprocal
        #     (ATen's scalar-divisor shortcut), so a row's scale can differ in its last bit and, rarely, one of its codes by one.
        #     Rows whose codes and scale agree must agree to the ulp; the others stay within one activation quantisation step.
        assert r["launches"] == (2 if name.startswith("int8") else 1), (name, r)
        if name.startswith("int8"):
            assert r["our_scale_is_amax_div_127"], (name, r)
            assert r["max_code_delta"] <= 1 and r["codes_differ"] <= 0.001 * r["rows"] * 4096, (name, r)
            assert r["max_ulp_rows_same"] <= 1, (name, r)
            assert r["max_abs_err_over_range"] <= 2e-2, (name, r)
        else:
            assert r["max_ulp"] <= 1 and r["frac_diff"] < 0.02, (name, r)


CHILD_ATTN = textwrap.dedent('''
    import json, os, re, sys
    sys.path.insert(0, os.getcwd())
    import torch
    from oracle.ref_loader import load_reference
    # one point of the reference's autotune space (read at import time, triton_atten.py:16-24): keeps this test to seconds
    os.environ.update(SDNQ_TRITON_ATTEN_BLOCK_SIZE_M_LIST="128", SDNQ_TRITON_ATTEN_BLOCK_SIZE_N_LIST="32", SDNQ_TRITON_ATTEN_NUM_WARPS_LIST="4",
                      SDNQ_TRITON_ATTEN_NUM_STAGES_LIST="2")
    sdnq = load_reference(SDNQ_DEVICE="cuda", SDNQ_USE_TORCH_COMPILE="0")
    from sdnq.kernels import triton_atten
    import triton
    triton.set_allocator(lambda size, align, stream: torch.empty(size, dtype=torch.int8, device="cuda"))      # harness set-up for the reference's device-side descriptors
    from sdnq_b200 import _lib
    md = open("INTEGRATION.md").read()
    blocks = re.findall(r"```python\\n(.*?)```", md, flags=re.S)
    stub = {}
    exec(compile(next(b for b in blocks if "sdnq/kernels/b200.py" in b).replace('"libsdnq_b200.so"', repr(_lib.LIB_PATH)), "INTEGRATION.md:b200.py", "exec"), stub)
    exec(compile(next(b for b in blocks if "sdnq/kernels/b200_atten.py" in b), "INTEGRATION.md:b200_atten.py", "exec"), stub)
    reference_launch = triton_atten.sdnq_atten_fwd
    out = {}
    cases = {"int8": (1, 4, 640, 640, 128, dict()), "fp8_causal": (2, 3, 384, 384, 64, dict(matmul_dtype="float8_e4m3fn", is_causal=True)),
             "bool_mask": (1, 2, 256, 320, 64, dict(mask=True)), "pv_int8": (1, 2, 512, 512, 128, dict(pv_matmul_dtype="int8"))}
    for name, (Z, H, QN, KN, HD, kw) in cases.items():
        g = torch.Generator().manual_seed(QN + HD)
        q = torch.randn(Z, H, QN, HD, generator=g).bfloat16().cuda()
        k = (torch.randn(Z, H, KN, HD, generator=g) + 0.5).bfloat16().cuda()
        v = torch.randn(Z, H, KN, HD, generator=g).bfloat16().cuda()
        kw = dict(kw)
        if kw.pop("mask", False):
            m = torch.rand(1, 1, QN, KN, generator=g) > 0.3
            m[..., 0] = True
            kw["attn_mask"] = m.cuda()
        with torch.no_grad():
            y_ref = triton_atten.sdnq_triton_atten(q, k, v, **kw).float()                 # the reference's own Triton kernel
            triton_atten.sdnq_atten_fwd = stub["sdnq_atten_fwd"]                          # triton_atten.py:338-386 re-pointed at libsdnq_b200.so
            _lib.launch_count(reset=True)
            y_stub = triton_atten.sdnq_triton_atten(q, k, v, **kw).float()                # same quantize_attn / get_attn_inputs, other kernel
            launches = _lib.launch_count()
            triton_atten.sdnq_atten_fwd = reference_launch
        scale = float(y_ref.abs().max())
        out[name] = {"launches": int(launches), "max_err_over_range": float((y_stub - y_ref).abs().max()) / scale,
                     "rel_l2": float((y_stub - y_ref).norm() / y_ref.norm()), "shape_ok": bool(y_stub.shape == y_ref.shape)}
    print("LEVEL_B_ATTN " + json.dumps(out))
''')


def test_reference_attention_over_the_c_abi_stub_matches_its_own_triton_kernel():
    """row f3 at level B: the reference's `sdnq_triton_atten` (its own smooth-K / quantize_attn / get_attn_inputs) with only the kernel
    launch `sdnq_atten_fwd` (kernels/triton_atten.py:338-386) replaced by the INTEGRATION.md stub over `sdnq_b200_attention`, against
    the same call on the reference's Triton kernel.  Tolerances as in tests/test_attention_gpu.py (bf16 outputs, different key-block sizes)."""
    sys.path.insert(0, ROOT)
    from oracle.ref_loader import reference_root
    if reference_root() is None:
        pytest.skip("oracle/_ref is not present (python oracle/build_ref.py in the authoring container)")
    p = subprocess.run([sys.executable, "-c", CHILD_ATTN], capture_output=True, text=True, cwd=ROOT, timeout=900)
    got = [ln for ln in p.stdout.splitlines() if ln.startswith("LEVEL_B_ATTN ")]
    if not got and "triton" in p.stderr.lower() and "error" in p.stderr.lower() and "sdnq_b200" not in p.stderr:
        pytest.skip("the reference's Triton attention does not run on this box: " + p.stderr.strip().splitlines()[-1][:300])
    assert p.returncode == 0 and got, (p.stdout[-1500:], p.stderr[-3000:])
    res = json.loads(got[-1][len("LEVEL_B_ATTN "):])
    print(json.dumps(res, indent=1))
    for name, r in res.items():
        assert r["shape_ok"] and r["launches"] == 2, (name, r)           # V^T pre-pass + K9
        loose = name.startswith("pv_")                                   # P's row scale is per key block: 128 here, 32 in the reference
        assert r["max_err_over_range"] <= (8e-2 if loose else 2e-2) and r["rel_l2"] <= (4e-2 if loose else 1e-2), (name, r)
