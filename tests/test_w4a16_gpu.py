"""K6 (gemm_w4a16.cu) on the GPU: the dequant-path Linear of 4-bit layers in one launch -- packed codes dequantised in the GEMM
prologue, SVD rank-r term as a second accumulate -- against
  * an exact-arithmetic construction (integer activations, power-of-two scales: every product and sum is exact, so any
    misplaced byte, swizzle slip or k-order error shows as a bit difference),
  * the numpy oracle of the reference's quantized_linear_forward (dequantise, f32-accumulated product, one rounding),
  * the reference-shaped path of this library (K3 dequant + library GEMM, SDNQ_B200_W4A16=0).
K6 is opt-in (SDNQ_B200_W4A16=1; dequantise-once + GEMM is faster at the BASELINE shapes, profiles/r02_w4a16.md); the tests switch it
on.  The reference's own outputs for this path (tests/golden/layer_c4_int4_g128_svd_dequant*.npz) are checked against both routes in
test_w4a16_reference_fixture below and tests/test_layers_gpu.py::test_forward_matches_reference_output."""
import copy

import numpy as np
import pytest
import torch

from oracle import sdnq_oracle as O
from tests.util import bf16_ulp_diff, oracle_layer_from_torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from sdnq_b200 import ops
    return ops


@pytest.mark.parametrize("M,N,K,group,rank,unsigned", [(128, 128, 64, 64, 0, False), (128, 128, 256, 128, 0, False), (200, 264, 640, 128, 0, True),
                                                       (77, 640, 2048, 128, 32, False), (1024, 1280, 1280, 128, 32, False), (160, 136, 320, 32, 16, True),
                                                       (300, 256, 512, 512, 64, False), (4096, 640, 640, 128, 32, False)])
def test_w4a16_exact_arithmetic(M, N, K, group, rank, unsigned):
    """x in {-2..2}, scales 2^-e, zero points k/4, svd factors sparse in {-1, 0, 1}: the f32 accumulators hold exact integers /
    dyadic rationals, so the kernel must reproduce the float64 result rounded once to bf16 -- bit for bit."""
    g = torch.Generator().manual_seed(M * 7 + N + K + rank)
    codes = torch.randint(0, 16, (N, K), generator=g)
    packed = (codes[:, 0::2] | (codes[:, 1::2] << 4)).to(torch.uint8).reshape(-1)          # packed_int/pack.py:273-276
    gpr = K // group
    scale = torch.pow(2.0, -torch.randint(1, 5, (N, gpr), generator=g).float())
    zp = (torch.randint(-8, 8, (N, gpr), generator=g).float() / 4) if unsigned else None
    x = torch.randint(-2, 3, (M, K), generator=g).float()
    bias = torch.randint(-8, 9, (N,), generator=g).float() / 2
    q = codes.double() - (0 if unsigned else 8)
    W = q * scale.double().repeat_interleave(group, dim=1)
    if unsigned:
        W = W + zp.double().repeat_interleave(group, dim=1)
    ref = x.double() @ W.t() + bias.double()
    down = up = None
    if rank:
        down = torch.zeros(rank, K)
        idx = torch.randint(0, K, (rank, 12), generator=g)
        down.scatter_(1, idx, torch.randint(0, 2, (rank, 12), generator=g).float() * 2 - 1)     # |x down^T| <= 24: exact in bf16
        up = torch.randint(-1, 2, (N, rank), generator=g).float()
        ref = ref + (x.double() @ down.double().t()) @ up.double().t()
    got = _ops().linear_w4a16(x.to(DEV, torch.bfloat16), packed.to(DEV), "uint4" if unsigned else "int4", scale.to(DEV), None if zp is None else zp.to(DEV),
                              N, K, bias=bias.to(DEV, torch.bfloat16), svd_down_rk=None if down is None else down.to(DEV, torch.bfloat16),
                              svd_up_nr=None if up is None else up.to(DEV, torch.bfloat16))
    want = ref.float().to(torch.bfloat16)
    assert got.shape == (M, N) and got.dtype == torch.bfloat16
    bad = (got.cpu() != want)
    assert not bool(bad.any()), f"{int(bad.sum())} of {M * N} outputs differ; first at {bad.nonzero()[0].tolist()}"


CFGS = {
    "int4_g128": dict(weights_dtype="int4", group_size=128),
    "int4_g128_svd32": dict(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32),
    "uint4_g64": dict(weights_dtype="uint4", group_size=64),
    "int4_rowwise": dict(weights_dtype="int4", group_size=-1),
    "uint4_g32_svd16": dict(weights_dtype="uint4", group_size=32, use_svd=True, svd_rank=16),
    "int4_g128_svd64": dict(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=64),
    "int4_g64_hadamard128": dict(weights_dtype="int4", group_size=64, use_hadamard=True, hadamard_group_size=128),
    "int4_g128_svd32_hadamard256": dict(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32, use_hadamard=True, hadamard_group_size=256),
}


@pytest.mark.parametrize("name", sorted(CFGS))
@pytest.mark.parametrize("M,N,K", [(48, 136, 256), (333, 640, 768)])
def test_w4a16_forward_vs_oracle_and_dequant_path(name, M, N, K, monkeypatch):
    from sdnq_b200 import SDNQConfig, _lib, sdnq_quantize_layer
    cfg = CFGS[name]
    torch.manual_seed(len(name) + M)
    lin = torch.nn.Linear(K, N, bias=True).to(torch.bfloat16)
    layer, _ = sdnq_quantize_layer(copy.deepcopy(lin), SDNQConfig(**cfg))
    assert layer.forward_func.__name__ == "quantized_linear_forward"
    x = torch.randn(M, K).to(torch.bfloat16)
    meta = {k: (list(v) if isinstance(v, (torch.Size, tuple)) else v) for k, v in layer.sdnq_dequantizer.__dict__.items() if k != "result_dtype"}
    ol = oracle_layer_from_torch({k: getattr(layer, k) for k in ("weight", "scale", "zero_point", "svd_up", "svd_down", "bias")}, meta)
    ref = O.linear_forward(ol, x.float().numpy())
    layer = layer.to(DEV)
    xd = x.to(DEV)
    monkeypatch.setenv("SDNQ_B200_W4A16", "1")
    _lib.launch_count(reset=True)
    y = layer(xd)
    n_launch = _lib.launch_count()
    assert n_launch == (2 if cfg.get("use_hadamard") else 1), f"expected the one-launch W4A16 kernel, saw {n_launch} launches"
    monkeypatch.setenv("SDNQ_B200_W4A16", "0")
    y_dq = layer(xd)                                      # K3 (+ un-rotate) + library GEMM: the reference's shape of this forward (the default)
    got = y.float().cpu().numpy()
    scale = float(np.abs(ref).max())
    err = np.abs(got - ref)
    # stated tolerance of every path with a 16-bit GEMM in it (tests/test_layers_gpu.py): 2e-2 of the range, 3e-3 rms
    assert err.max() <= 2e-2 * scale and np.sqrt((err ** 2).mean()) <= 3e-3 * scale, (err.max(), scale)
    e2 = (y.float() - y_dq.float()).abs()
    assert float(e2.max()) <= 2e-2 * scale and float(e2.pow(2).mean().sqrt()) <= 3e-3 * scale
    if not cfg.get("use_svd") and not cfg.get("use_hadamard"):
        # same bf16 weights on both sides, f32 accumulation: only the summation order differs
        du = bf16_ulp_diff(y, y_dq)
        assert float((du > 1).float().mean()) < 0.01 and float(e2.max()) <= 2.0 ** -7 * scale
        du_o = bf16_ulp_diff(y.cpu(), torch.from_numpy(ref).to(torch.bfloat16))
        assert float((du_o > 1).float().mean()) < 0.01


def test_w4a16_f16_activations_and_fallbacks(monkeypatch):
    """f16 activations run the f16-format MMA; shapes the kernel does not cover (K % 64 != 0) keep the K3 + GEMM path."""
    from sdnq_b200 import SDNQConfig, _lib, sdnq_quantize_layer
    monkeypatch.setenv("SDNQ_B200_W4A16", "1")
    torch.manual_seed(5)
    lin = torch.nn.Linear(512, 256, bias=True).to(torch.float16)
    layer, _ = sdnq_quantize_layer(copy.deepcopy(lin), SDNQConfig(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32))
    layer = layer.to(DEV)
    x = torch.randn(96, 512, dtype=torch.float16, device=DEV)
    _lib.launch_count(reset=True)
    y = layer(x)
    assert _lib.launch_count() == 1 and y.dtype == torch.float16
    W = layer.sdnq_dequantizer(layer.weight, layer.scale, svd_up=layer.svd_up, svd_down=layer.svd_down)
    ref = torch.nn.functional.linear(x.float(), W.float(), layer.bias.float())
    scale = float(ref.abs().max())
    assert float((y.float() - ref).abs().max()) <= 1e-2 * scale
    lin2 = torch.nn.Linear(96, 64, bias=True).to(torch.bfloat16)          # K = 96: not a multiple of 64
    layer2, _ = sdnq_quantize_layer(copy.deepcopy(lin2), SDNQConfig(weights_dtype="int4", group_size=32, minimum_allowed_numel=1))
    layer2 = layer2.to(DEV)
    y2 = layer2(torch.randn(40, 96, dtype=torch.bfloat16, device=DEV))
    assert y2.shape == (40, 64) and bool(torch.isfinite(y2).all())


@pytest.mark.parametrize("fixture", ["layer_c4_int4_g128_svd_dequant.npz", "layer_uint4_auto_dequant.npz"])
def test_w4a16_reference_fixture(fixture, monkeypatch):
    """the reference's own output for dequant-path 4-bit layers (tests/golden) through K6"""
    import os

    from sdnq_b200 import _lib
    from tests.test_layers_gpu import build_layer
    from tests.util import GOLDEN, np_to_torch
    monkeypatch.setenv("SDNQ_B200_W4A16", "1")
    layer, t, z, meta = build_layer(os.path.join(GOLDEN, fixture))
    _lib.launch_count(reset=True)
    y = layer(t["x"].to(DEV))
    assert _lib.launch_count() == 1
    yref = np_to_torch(z["y"], "bfloat16", DEV)
    scale = float(yref.float().abs().max())
    err = (y.float() - yref.float()).abs()
    assert float(err.max()) <= 2e-2 * scale and float(err.pow(2).mean().sqrt()) <= 3e-3 * scale
