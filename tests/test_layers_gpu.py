"""Layer-level parity on the GPU: SDNQLinear.forward through the public surface (SDNQConfig -> sdnq_quantize_layer ->
forward_func -> C ABI kernels) against the reference outputs recorded in tests/golden/, plus size-independent properties at
the BASELINE.json shapes."""
import copy

import numpy as np
import pytest
import torch

from oracle import sdnq_oracle as O
from tests.util import LAYER_FILES, LAYER_IDS, bf16_ulp_diff, fixture_tensors, np_to_torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def build_layer(path):
    """quantise w_orig with our host code, then pin the stored tensors to the reference's (bit-identical anyway unless SVD)."""
    from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
    t, z, meta = fixture_tensors(path)
    lin = torch.nn.Linear(meta["K"], meta["N"], bias=t["bias"] is not None).to(torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(t["w_orig"])
        if t["bias"] is not None:
            lin.bias.copy_(t["bias"])
    layer, _ = sdnq_quantize_layer(copy.deepcopy(lin), SDNQConfig(**meta["config"]))
    for key in ("weight", "scale", "zero_point", "svd_up", "svd_down"):
        if t[key] is not None:
            setattr(layer, key, torch.nn.Parameter(t[key], requires_grad=False))
    return layer.to(DEV), t, z, meta


@pytest.mark.parametrize("path", LAYER_FILES, ids=LAYER_IDS)
def test_forward_matches_reference_output(path):
    layer, t, z, meta = build_layer(path)
    d = meta["dequantizer"]
    y = layer(t["x"].to(DEV))
    yref = np_to_torch(z["y"], "bfloat16", DEV)
    assert y.shape == yref.shape and y.dtype == torch.bfloat16
    finite = torch.isfinite(yref)
    assert torch.equal(finite, torch.isfinite(y))
    is_mm = d["use_quantized_matmul"] and meta["M"] >= 32
    scale = float(yref[finite].float().abs().max())
    err = (y.float() - yref.float())[finite].abs()
    if is_mm and not d["use_hadamard"] and t["svd_up"] is None:
        # integer / fp8 contraction is exact, activation codes are bit-exact: only the f32 epilogue rounding can move 1 bf16 ulp
        du = bf16_ulp_diff(y, yref)[finite]
        assert int(du.max()) <= 1, f"max {int(du.max())} ulp"
        assert float((du > 0).float().mean()) < 0.02
    else:
        # tolerance for paths with a bf16 GEMM / Hadamard / SVD in them: 2e-2 of the output range max, 3e-3 rms (bf16 has 8 bits)
        assert float(err.max()) <= 2e-2 * scale and float(err.pow(2).mean().sqrt()) <= 3e-3 * scale


@pytest.mark.parametrize("path", [p for p in LAYER_FILES if "c2_int8_w8a8." in p or "uint4_auto_dequant" in p or "int8_svd_w8a8" in p],
                         ids=["int8_w8a8", "int8_svd_w8a8", "uint4_dequant"])
def test_dequantize_restores_module(path):
    layer, t, z, meta = build_layer(path)
    x = t["x"].to(DEV)
    y_q = layer(x)
    dense = layer.dequantize()
    assert type(dense) is torch.nn.Linear and not hasattr(dense, "sdnq_dequantizer")
    wref = np_to_torch(z["w_dequant"], "bfloat16", DEV)
    assert dense.weight.shape == wref.shape
    assert float((dense.weight.float() - wref.float()).abs().max()) <= 2.0 ** -7 * float(wref.float().abs().max())
    y_d = dense(x)
    assert float((y_d.float() - y_q.float()).abs().max()) <= 0.05 * float(y_q.float().abs().max())


def _model(K=256, N=128):
    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(K, N)
            self.b = torch.nn.Linear(N, K)

        def forward(self, x):
            return self.b(torch.nn.functional.gelu(self.a(x)))
    torch.manual_seed(4)
    return Net().to(torch.bfloat16)


@pytest.mark.parametrize("cfg", [dict(weights_dtype="int8"), dict(weights_dtype="uint8"), dict(weights_dtype="float8_e4m3fn", use_hadamard=True),
                                 dict(weights_dtype="int8", use_svd=True, svd_rank=8), dict(weights_dtype="int4", group_size=64)],
                         ids=["int8", "uint8", "fp8_hadamard", "int8_svd", "int4_g64"])
def test_apply_sdnq_options_flips_matmul_in_place(cfg):
    """quantise with the matmul off, flip it on with apply_sdnq_options_to_model, and compare with a model quantised with the
    matmul on from the start: same stored bytes, same outputs (reference loader.py:286-301)."""
    from sdnq_b200 import apply_sdnq_options_to_model, sdnq_post_load_quant
    base = _model()
    torch.manual_seed(1)
    off = sdnq_post_load_quant(copy.deepcopy(base), use_quantized_matmul=False, add_skip_keys=False, **cfg).to(DEV)
    torch.manual_seed(1)
    on = sdnq_post_load_quant(copy.deepcopy(base), use_quantized_matmul=True, add_skip_keys=False, **cfg).to(DEV)
    x = torch.randn(64, 256, dtype=torch.bfloat16, device=DEV)
    y_off = off(x)
    flipped = apply_sdnq_options_to_model(off, use_quantized_matmul=True)
    assert flipped.a.forward_func.__name__ == on.a.forward_func.__name__
    assert flipped.a.sdnq_dequantizer.use_quantized_matmul is True
    if not cfg.get("use_svd"):
        for key in ("weight", "scale", "zero_point"):
            u, v = getattr(flipped.a, key), getattr(on.a, key)
            assert (u is None) == (v is None)
            if u is not None:
                assert u.shape == v.shape and u.stride() == v.stride()
                assert torch.equal(u.view(torch.uint8) if u.dtype == torch.float8_e4m3fn else u, v.view(torch.uint8) if v.dtype == torch.float8_e4m3fn else v)
        assert torch.equal(flipped(x), on(x))
    y_on = flipped(x)
    assert float((y_on.float() - y_off.float()).abs().max()) <= 0.08 * float(y_off.float().abs().max())
    back = apply_sdnq_options_to_model(flipped, use_quantized_matmul=False)
    assert back.a.forward_func.__name__ == "quantized_linear_forward"
    assert torch.equal(back(x), y_off)


def test_cpu_tensor_fails_loudly():
    from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
    from sdnq_b200._lib import SDNQKernelError
    layer, _ = sdnq_quantize_layer(torch.nn.Linear(64, 64).to(torch.bfloat16), SDNQConfig(weights_dtype="int8", use_quantized_matmul=True, minimum_allowed_numel=1))
    with pytest.raises(SDNQKernelError):
        layer(torch.randn(40, 64, dtype=torch.bfloat16))


# --------------------------------------------------------------------------------------- BASELINE-size properties
SDXL = [(4096, 640, 640), (4096, 5120, 640), (1024, 1280, 5120), (77, 1280, 2048)]
FLUX = [(16384, 3072, 3072), (2048, 12288, 3072)]


@pytest.mark.parametrize("M,N,K", SDXL + FLUX)
@pytest.mark.parametrize("wd", ["int8", "float8_e4m3fn"])
def test_full_size_properties(M, N, K, wd):
    """at the real SD-XL / FLUX Linear shapes: (1) row independence -- permuting the rows of x permutes the rows of y bit-exactly;
    (2) power-of-two linearity -- y(4x) - bias == 4 (y(x) - bias) up to the output rounding (codes identical, scales x4);
    (3) agreement with the oracle on a random sample of rows (oracle evaluated on those rows only)."""
    from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
    torch.manual_seed(M + N + K)
    lin = torch.nn.Linear(K, N, bias=True).to(torch.bfloat16)
    layer, _ = sdnq_quantize_layer(copy.deepcopy(lin), SDNQConfig(weights_dtype=wd, use_quantized_matmul=True, use_hadamard=(wd != "int8")))
    layer = layer.to(DEV)
    x = torch.randn(M, K, dtype=torch.bfloat16, device=DEV)
    y = layer(x)
    perm = torch.randperm(M, device=DEV)
    assert torch.equal(layer(x[perm]), y[perm])
    bias = layer.bias
    layer.bias = None
    y0, y4 = layer(x), layer(x * 4)
    layer.bias = bias
    assert torch.equal(y4, y0 * 4)
    rows = torch.randperm(M)[:48].sort().values
    meta = {k: (list(v) if isinstance(v, torch.Size) else v) for k, v in layer.sdnq_dequantizer.__dict__.items() if k != "result_dtype"}
    ol = O.Layer(layer.weight.detach().float().cpu().numpy() if wd != "int8" else layer.weight.detach().cpu().numpy(),
                 layer.scale.detach().cpu().numpy(), bias=layer.bias.detach().float().cpu().numpy(), **meta)
    ref = O.linear_forward(ol, x[rows.to(DEV)].float().cpu().numpy())
    got = y[rows.to(DEV)].float().cpu().numpy()
    scale = np.abs(ref).max()
    if wd == "int8":
        du = bf16_ulp_diff(torch.from_numpy(got).to(torch.bfloat16), torch.from_numpy(ref).to(torch.bfloat16))
        assert int(du.max()) <= 1
    else:
        assert np.abs(got - ref).max() <= 2e-2 * scale and np.sqrt(np.mean((got - ref) ** 2)) <= 3e-3 * scale


@pytest.mark.parametrize("cfg", [dict(weights_dtype="int8"), dict(weights_dtype="uint8"), dict(weights_dtype="float8_e4m3fn"),
                                 dict(weights_dtype="float8_e4m3fn", use_hadamard=True, hadamard_group_size=256),
                                 dict(weights_dtype="int8", use_hadamard=True, hadamard_group_size=128)],
                         ids=["int8", "uint8", "fp8", "fp8_hadamard256", "int8_hadamard128"])
@pytest.mark.parametrize("M", [1, 4, 31])
def test_small_m_forward_gemv_vs_dequant_path(cfg, M, monkeypatch):
    """rows < 32 of a W8A8 layer: the K5 GEMV (default) against the reference-shaped dequantise + bf16 GEMM path
    (SDNQ_B200_SMALL_M_GEMV=0), which is itself pinned to the reference by the `small_m` fixture.  Tolerance = the one stated for
    every path with a 16-bit GEMM in it (the dequant path rounds each weight to bf16, the GEMV does not)."""
    from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
    torch.manual_seed(7 + M)
    lin = torch.nn.Linear(768, 1536, bias=True).to(torch.bfloat16)
    layer, _ = sdnq_quantize_layer(copy.deepcopy(lin), SDNQConfig(use_quantized_matmul=True, **cfg))
    layer = layer.to(DEV)
    x = torch.randn(M, 768, dtype=torch.bfloat16, device=DEV)
    from sdnq_b200 import _lib
    _lib.launch_count(reset=True)
    y = layer(x)
    n_gemv = _lib.launch_count()
    monkeypatch.setenv("SDNQ_B200_SMALL_M_GEMV", "0")
    y_ref = layer(x)
    monkeypatch.delenv("SDNQ_B200_SMALL_M_GEMV")
    assert n_gemv == (2 if cfg.get("use_hadamard") else 1)
    dense = lin.to(DEV)(x)
    scale = float(y_ref.float().abs().max())
    err = (y.float() - y_ref.float()).abs()
    assert float(err.max()) <= 2e-2 * scale and float(err.pow(2).mean().sqrt()) <= 3e-3 * scale
    # and both are the same distance from the unquantised layer (quantisation error dominates)
    e1, e2 = (y.float() - dense.float()).pow(2).mean().sqrt(), (y_ref.float() - dense.float()).pow(2).mean().sqrt()
    assert float(e1) <= 1.1 * float(e2) + 1e-3 * scale


@pytest.mark.parametrize("name", ["uint4_dynamic", "int3_dynamic_tight", "uint4_dynamic_w8a8"])
def test_dynamic_quantization_picks_the_reference_dtypes(name):
    """use_dynamic_quantization (quantizer.py:280-419): walk weights_dtype_order until the normalised MSE of quantise -> dequantise
    (through the K3 kernel here) drops below the threshold.  The dtype chosen for every layer, the wrapper classes, forward
    functions, dequantizer metadata and stored shapes equal the reference's run on the same toy model
    (tests/golden/model_dynamic.json), and the quantised model still tracks the dense one."""
    import json
    import os
    import sys
    from tests.util import GOLDEN
    sys.path.insert(0, GOLDEN)
    import toy_model
    from sdnq_b200 import sdnq_post_load_quant
    ref = json.load(open(os.path.join(GOLDEN, "model_dynamic.json")))[name]
    dense = toy_model.build().to(DEV)
    model = sdnq_post_load_quant(toy_model.build().to(DEV), **ref["config"])
    qc = model.quantization_config
    assert {k: sorted(v) for k, v in qc.modules_dtype_dict.items()} == ref["modules_dtype_dict"]
    assert sorted(qc.modules_to_not_convert) == ref["modules_to_not_convert"]
    assert sorted(qc.modules_to_not_use_matmul) == ref["modules_to_not_use_matmul"]
    mods = dict(model.named_modules())
    for mname, e in ref["modules"].items():
        m = mods[mname]
        assert type(m).__name__ == e["class"], mname
        if "forward_func" in e:
            assert m.forward_func.__name__ == e["forward_func"], mname
            for k in ("weights_dtype", "group_size", "quantized_matmul_dtype", "use_quantized_matmul", "re_quantize_for_matmul", "quantized_weight_shape"):
                got = getattr(m.sdnq_dequantizer, k)
                got = list(got) if isinstance(got, (torch.Size, tuple)) else got
                assert got == e["dequantizer"][k], (mname, k, got, e["dequantizer"][k])
    sd = model.state_dict()
    for key, e in ref["tensors"].items():
        assert key in sd and list(sd[key].shape) == e["shape"] and str(sd[key].dtype).replace("torch.", "") == e["dtype"], key
    x = torch.randn(40, 128, dtype=torch.bfloat16, device=DEV)
    y, yd = model.mid[0].ff(x), dense.mid[0].ff(x)
    assert float((y.float() - yd.float()).pow(2).mean().sqrt()) <= 0.1 * float(yd.float().pow(2).mean().sqrt())


@pytest.mark.parametrize("cfg", [dict(weights_dtype="int8", use_quantized_matmul=True), dict(weights_dtype="uint8", use_quantized_matmul=True),
                                 dict(weights_dtype="uint4"), dict(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32),
                                 dict(weights_dtype="int6", use_quantized_matmul=True), dict(weights_dtype="int5", group_size=32),
                                 dict(weights_dtype="uint3", group_size=64, dequantize_fp32=False), dict(weights_dtype="int2", group_size=16),
                                 dict(weights_dtype="int8", use_quantized_matmul=True, use_hadamard=True, hadamard_group_size=128),
                                 dict(weights_dtype="float8_e4m3fn", use_quantized_matmul=True), dict(weights_dtype="float8_e5m2", group_size=64),
                                 dict(weights_dtype="float6_e3m2fn", use_quantized_matmul=True), dict(weights_dtype="float4_e2m1fn", group_size=32),
                                 dict(weights_dtype="float8_e4m3fn_sdnq"), dict(weights_dtype="float7_e3m4fnu", group_size=128)],
                         ids=lambda c: "_".join(str(v) for v in c.values()))
def test_quantising_on_the_gpu_stores_what_the_cpu_path_stores(cfg, monkeypatch):
    """sdnq_quantize_layer on a CUDA weight goes through K8 (scale + round + clamp + pack in one kernel); the stored tensors are the
    ones the host arithmetic (the reference's, on the CPU) produces.  SVD layers: the factors come from a randomised decomposition, so the
    comparison is made on the kernel's input (the residual weight) by running the eager ops on the same device."""
    from sdnq_b200 import SDNQConfig, _lib, sdnq_quantize_layer
    torch.manual_seed(3)
    lin = torch.nn.Linear(640, 384).to(torch.bfloat16)
    if cfg.get("use_svd") or cfg.get("use_hadamard"):
        ref_dev = DEV              # same device, eager ops vs kernel
    else:
        ref_dev = "cpu"
    monkeypatch.setenv("SDNQ_B200_QUANT_KERNEL", "0")
    torch.manual_seed(11)
    ref, _ = sdnq_quantize_layer(copy.deepcopy(lin).to(ref_dev), SDNQConfig(**cfg))
    monkeypatch.setenv("SDNQ_B200_QUANT_KERNEL", "1")
    _lib.launch_count(reset=True)
    torch.manual_seed(11)
    got, _ = sdnq_quantize_layer(copy.deepcopy(lin).to(DEV), SDNQConfig(**cfg))
    assert _lib.launch_count() >= 1
    for name in ("weight", "scale", "zero_point"):
        a, b = getattr(got, name), getattr(ref, name)
        assert (a is None) == (b is None), name
        if a is None:
            continue
        assert a.dtype == b.dtype and a.shape == b.shape and a.stride() == b.stride(), (name, a.dtype, b.dtype, a.shape, b.shape, a.stride(), b.stride())
        if ref_dev == "cpu":
            if a.element_size() == 1:
                a, b = a.view(torch.uint8), b.view(torch.uint8)          # (float8 tensors: compare the bytes)
            assert torch.equal(a.cpu(), b.cpu()), name
        else:
            # eager CUDA divides by a scalar through its reciprocal; the kernel divides: scales may differ in the last bit, codes by one
            if a.is_floating_point():
                assert torch.allclose(a.float(), b.float(), rtol=2e-7 if a.dtype == torch.float32 else 0, atol=0), name
            else:
                diff = (a.view(torch.uint8).int() - b.view(torch.uint8).int()).abs()
                assert float((diff != 0).float().mean()) < 5e-3, name
    x = torch.randn(64, 640, device=DEV, dtype=torch.bfloat16)
    assert torch.isfinite(got(x)).all()


@pytest.mark.parametrize("cfg", [dict(weights_dtype="int8"), dict(weights_dtype="uint4", group_size=32), dict(weights_dtype="int5", group_size=64),
                                 dict(weights_dtype="int4", group_size=64, use_svd=True, svd_rank=16), dict(weights_dtype="int8", use_hadamard=True, hadamard_group_size=64),
                                 dict(weights_dtype="float6_e3m2fn"), dict(weights_dtype="uint4", use_codebook=True, group_size=32)],
                         ids=lambda c: "_".join(str(v) for v in c.values()))
def test_quantized_embedding_forward(cfg):
    """quant_embedding=True (reference layers/embedding/forward.py:14-104: index the unpacked table, dequantise the selected rows):
    the gather + dequant kernel gives exactly the rows of the fully dequantised table (the K3 arithmetic the reference fixtures pin),
    times scalar_embed_scale."""
    from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
    torch.manual_seed(4)
    emb = torch.nn.Embedding(1000, 256).to(torch.bfloat16)
    layer, _ = sdnq_quantize_layer(copy.deepcopy(emb), SDNQConfig(quant_embedding=True, **cfg))
    assert type(layer).__name__ == "SDNQEmbedding" and layer.forward_func.__name__ == "quantized_embedding_forward"
    layer = layer.to(DEV)
    ids = torch.randint(0, 1000, (3, 77), device=DEV)
    ids[0, 0], ids[0, 1] = 0, 999
    d = layer.sdnq_dequantizer
    table = d(layer.weight, layer.scale, zero_point=layer.zero_point, svd_up=layer.svd_up, svd_down=layer.svd_down)
    got = layer(ids)
    assert got.shape == (3, 77, 256) and got.dtype == torch.bfloat16 and got.is_contiguous()
    assert torch.equal(got, table[ids])
    err = (got.float() - emb.weight.to(DEV).float()[ids]).abs().max()
    assert float(err) < 0.25 * float(emb.weight.float().abs().max())                      # it is the embedding it was quantised from
    layer.scalar_embed_scale = 16.0                                                         # Gemma-style scaled embeddings
    assert torch.equal(layer(ids), table[ids].mul_(16.0))
