"""K5p on the GPU: the small-M Linear straight from the stored (packed / group-wise) weight -- the default for rows < 32 of every
packed / group-wise layer (SDNQ_B200_SMALL_M_PACKED=0 restores dequantise + GEMM).

Against the reference itself: the `*_small_m` layer fixtures of tests/golden/ run through it in
tests/test_layers_gpu.py::test_forward_matches_reference_output.  Here: a wider format x shape sweep against the reference-shaped
dequantise + bf16 GEMM path (which those fixtures pin to the reference) and against the oracle directly.
(First hardware run: the round-1 driver box, 36 / 36 XPASS.)"""
import copy

import pytest
import torch

from tests.util import bf16_ulp_diff

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("cfg", [dict(weights_dtype="int4", group_size=128), dict(weights_dtype="uint4"), dict(weights_dtype="int2", group_size=16),
                                 dict(weights_dtype="float6_e3m2fn", group_size=32), dict(weights_dtype="int5", group_size=-1),
                                 dict(weights_dtype="int4", group_size=-1, use_quantized_matmul=True),
                                 dict(weights_dtype="int8", group_size=128, use_quantized_matmul=True),
                                 dict(weights_dtype="uint4", use_quantized_matmul=True, use_hadamard=True, hadamard_group_size=256),
                                 dict(weights_dtype="int4", group_size=64, use_hadamard=True, hadamard_group_size=128),
                                 dict(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32),
                                 dict(weights_dtype="uint4", use_svd=True, svd_rank=16, use_quantized_matmul=True),
                                 dict(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32, use_hadamard=True, hadamard_group_size=256)],
                         ids=["int4_g128", "uint4_auto", "int2_g16", "float6_g32", "int5_rowwise", "int4_rowwise_w8a8", "int8_g128_w8a8",
                              "uint4_hadamard_w8a8", "int4_g64_hadamard128", "int4_g128_svd32", "uint4_svd16_w8a8", "int4_g128_svd32_hadamard256"])
@pytest.mark.parametrize("M", [1, 4, 31])
def test_small_m_packed_forward_vs_dequant_path(cfg, M, monkeypatch):
    """rows < 32 of a layer stored packed / group-wise: K5p (the default; reads the stored bytes once) against the
    reference-shaped dequantise + bf16 GEMM path (SDNQ_B200_SMALL_M_PACKED=0).  K5p multiplies by exactly the bf16 weights the dequant kernel would write, so
    only the f32 accumulation order differs from the library GEMM."""
    from sdnq_b200 import SDNQConfig, _lib, sdnq_quantize_layer
    torch.manual_seed(11 + M)
    # W8A8 needs N % 16 == 0 (utils.py:93-98), otherwise the layer silently takes the dequant forward; the others get a ragged last tile
    w8a8 = bool(cfg.get("use_quantized_matmul"))
    lin = torch.nn.Linear(768, 1552 if w8a8 else 1544, bias=True).to(torch.bfloat16)
    layer, _ = sdnq_quantize_layer(copy.deepcopy(lin), SDNQConfig(**cfg))
    assert layer.forward_func.__name__.endswith("_matmul") == w8a8, layer.forward_func.__name__
    layer = layer.to(DEV)
    x = torch.randn(M, 768, dtype=torch.bfloat16, device=DEV)
    monkeypatch.setenv("SDNQ_B200_SMALL_M_PACKED", "1")
    _lib.launch_count(reset=True)
    y = layer(x)
    n_launch = _lib.launch_count()
    monkeypatch.setenv("SDNQ_B200_SMALL_M_PACKED", "0")
    y_ref = layer(x)
    assert n_launch == (2 if cfg.get("use_hadamard") else 1), n_launch              # (the SVD term's two skinny GEMMs are library calls)
    assert y.shape == y_ref.shape and y.dtype == y_ref.dtype and bool(torch.isfinite(y).all())
    scale = float(y_ref.float().abs().max())
    err = (y.float() - y_ref.float()).abs()
    assert float(err.max()) <= 2e-2 * scale and float(err.pow(2).mean().sqrt()) <= 3e-3 * scale
    if not cfg.get("use_hadamard") and not cfg.get("use_svd"):      # same bf16 weights, f32 accumulation: at most the last bf16 bit of an output moves
        assert int(bf16_ulp_diff(y, y_ref).max()) <= 2 or float(err.max()) <= 2.0 ** -7 * scale


@pytest.mark.parametrize("cfg", [dict(weights_dtype="int4", group_size=128), dict(weights_dtype="uint4", group_size=64), dict(weights_dtype="int3", group_size=32),
                                 dict(weights_dtype="float5_e2m2fn", group_size=-1), dict(weights_dtype="uint7", group_size=64),
                                 dict(weights_dtype="int6", group_size=-1, use_quantized_matmul=True)],
                         ids=["int4_g128", "uint4_g64", "int3_g32", "float5_rowwise", "uint7_g64", "int6_rowwise_w8a8"])
@pytest.mark.parametrize("M", [2, 17])
def test_small_m_packed_forward_vs_oracle(cfg, M):
    """K5p against the numpy oracle of the reference's rows < 32 branch (dequantise to bf16, f32-accumulated product, one
    rounding): the kernel multiplies by exactly those bf16 weights, so the two differ by f32 summation order only."""
    import numpy as np

    from oracle import sdnq_oracle as O
    from sdnq_b200 import SDNQConfig, _lib, sdnq_quantize_layer
    from tests.util import oracle_layer_from_torch
    torch.manual_seed(23 + M)
    w8a8 = bool(cfg.get("use_quantized_matmul"))
    lin = torch.nn.Linear(512, 272, bias=True).to(torch.bfloat16)
    layer, _ = sdnq_quantize_layer(copy.deepcopy(lin), SDNQConfig(**cfg))
    assert layer.forward_func.__name__.endswith("_matmul") == w8a8
    x = torch.randn(M, 512, dtype=torch.bfloat16)
    meta = {k: (list(v) if isinstance(v, (torch.Size, tuple)) else v) for k, v in layer.sdnq_dequantizer.__dict__.items() if k != "result_dtype"}
    ol = oracle_layer_from_torch({k: getattr(layer, k) for k in ("weight", "scale", "zero_point", "svd_up", "svd_down", "bias")}, meta)
    ref = O.linear_forward(ol, x.float().numpy())
    layer = layer.to(DEV)
    _lib.launch_count(reset=True)
    got = layer(x.to(DEV))
    assert _lib.launch_count() == 1, "rows < 32 of a packed layer must be one K5p launch"
    got = got.float().cpu().numpy()
    scale = float(np.abs(ref).max())
    du = bf16_ulp_diff(torch.from_numpy(got).to(torch.bfloat16), torch.from_numpy(ref).to(torch.bfloat16))
    assert np.abs(got - ref).max() <= 2.0 ** -7 * scale and float((du > 1).float().mean()) < 0.01, (np.abs(got - ref).max(), scale)
