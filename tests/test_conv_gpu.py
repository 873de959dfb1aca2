"""Convolution path on the GPU against the reference-generated fixtures tests/golden/conv_*.npz:
conv-weight dequant (broadcast scales), the im2col-gather activation quantiser, and SDNQConv*.forward through the public
surface (SDNQConfig -> sdnq_quantize_layer -> forward_func -> C ABI kernels)."""
import numpy as np
import pytest
import torch

from oracle import sdnq_oracle as O
from tests.util import CONV_FILES, CONV_IDS, bf16_ulp_diff, build_conv_layer, fixture_tensors, np_to_torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ops():
    from sdnq_b200 import ops as _ops
    return _ops


def build_conv(path):
    layer, t, z, meta = build_conv_layer(path)
    return layer.to(DEV), t, z, meta


@pytest.mark.parametrize("path", CONV_FILES, ids=CONV_IDS)
def test_conv_forward_matches_reference_output(path):
    layer, t, z, meta = build_conv(path)
    d = meta["dequantizer"]
    x = t["x"].to(DEV)
    before = ops()._lib.launch_count(reset=True)
    y = layer(x)
    assert ops()._lib.launch_count() >= 1, "no sdnq_b200 kernel was launched"
    yref = np_to_torch(z["y"], "bfloat16", DEV)
    assert y.shape == yref.shape and y.dtype == torch.bfloat16 and y.is_contiguous()
    is_mm = d["use_quantized_matmul"] and x.numel() / x.shape[2] >= 32
    scale = float(yref.float().abs().max())
    err = (y.float() - yref.float()).abs()
    if is_mm and not d["use_hadamard"] and t["svd_up"] is None:
        du = bf16_ulp_diff(y, yref)      # exact contraction, bit-exact activation codes: only the f32 epilogue can move one bf16 ulp
        assert int(du.max()) <= 1 and float((du > 0).float().mean()) < 0.02
    else:
        assert float(err.max()) <= 2e-2 * scale and float(err.pow(2).mean().sqrt()) <= 3e-3 * scale


@pytest.mark.parametrize("path", [p for p in CONV_FILES if "w_dequant" in np.load(p).files],
                         ids=[i for p, i in zip(CONV_FILES, CONV_IDS) if "w_dequant" in np.load(p).files])
def test_conv_weight_dequant_matches_reference(path):
    layer, t, z, meta = build_conv(path)
    d = layer.sdnq_dequantizer
    W = d(layer.weight, layer.scale, zero_point=layer.zero_point, svd_up=layer.svd_up, svd_down=layer.svd_down,
          skip_quantized_matmul=d.use_quantized_matmul)
    Wref = np_to_torch(z["w_dequant"], "bfloat16", DEV)
    assert W.shape == Wref.shape and W.dtype == torch.bfloat16
    du = bf16_ulp_diff(W, Wref)
    if t["svd_up"] is None:
        assert int(du.max()) == 0, f"max {int(du.max())} ulp"
    else:   # the SVD product comes from a library bf16 GEMM on both sides; accumulation order may differ
        assert int(du.max()) <= 1 and float((du > 0).float().mean()) < 0.02


@pytest.mark.parametrize("path", [p for p in CONV_FILES if "mm_xq" in np.load(p).files],
                         ids=[i for p, i in zip(CONV_FILES, CONV_IDS) if "mm_xq" in np.load(p).files])
def test_conv_act_quant_matches_reference(path):
    """the im2col gather + row quantiser against the reference's F.unfold + quantize_*_mm_input: bit-exact codes and scales."""
    t, z, meta = fixture_tensors(path)
    kw, d = meta["module_kwargs"], meta["dequantizer"]
    x = t["x"].to(DEV)
    nd = x.ndim - 2
    tup = lambda v: O._tuple_n(tuple(v) if isinstance(v, list) else v, nd)  # noqa: E731
    k, s, p, dl = tup(kw["kernel_size"]), tup(kw.get("stride", 1)), tup(kw.get("padding", 0)), tup(kw.get("dilation", 1))
    x4 = x
    if nd == 1:
        x4, k, s, p, dl = x.unsqueeze(2), (1, k[0]), (1, s[0]), (0, p[0]), (1, dl[0])
    if kw.get("padding_mode", "zeros") != "zeros":       # process_conv_input pads first, then unfolds without padding
        x4 = torch.nn.functional.pad(x4, [v for pi in reversed(p) for v in (pi, pi)], mode=kw["padding_mode"])
        p = (0, 0)
    hg = d["hadamard_group_size"] if d["use_hadamard"] else 0
    mm = d["quantized_matmul_dtype"]
    xq, sx, zx, rowsum, x_rot, _ = ops().conv_act_quant(x4, k, s, p, dl, mm, hadamard_group=hg, want_rowsum=True, want_x_rot=True)
    cols_ref = np_to_torch(z["cols"], "bfloat16", DEV)
    if hg:
        assert float((bf16_ulp_diff(x_rot, cols_ref) > 1).float().mean()) < 2e-3
    else:
        assert torch.equal(x_rot, cols_ref), "the gathered im2col rows differ from F.unfold"
    ref_xq = z["mm_xq"]
    got = xq.view(torch.uint8).cpu().numpy() if xq.dtype == torch.float8_e4m3fn else xq.cpu().numpy()
    if hg:
        assert float(np.mean(got.reshape(ref_xq.shape) != ref_xq.view(got.dtype))) < 5e-3
    else:
        if xq.dtype == torch.float8_e4m3fn:
            assert np.array_equal(O.from_e4m3fn_bits(got.reshape(ref_xq.shape)), O.from_e4m3fn_bits(ref_xq))
        else:
            assert np.array_equal(got.reshape(ref_xq.shape), ref_xq)
        assert np.array_equal(sx.cpu().numpy(), z["mm_sx"].reshape(-1))
        if "mm_zx" in z.files:
            assert np.array_equal(zx.cpu().numpy(), z["mm_zx"].reshape(-1))
        assert np.array_equal(rowsum.cpu().numpy(), xq.cpu().to(torch.int32).sum(-1).numpy()) if xq.dtype == torch.int8 else True


@pytest.mark.parametrize("C,H,W,N,k,s,p", [(320, 64, 64, 320, 3, 1, 1), (640, 32, 32, 640, 3, 1, 1), (320, 64, 64, 320, 3, 2, 1), (1280, 16, 16, 1280, 1, 1, 0)])
def test_conv_sdxl_shapes_against_unfold(C, H, W, N, k, s, p):
    """SD-XL ResNet / down-sampler conv shapes: the gather quantiser equals act_quant over a materialised F.unfold (bit-exact)
    and the W8A8 conv equals the W8A8 Linear on those columns (bit-exact: same kernels, same operands)."""
    torch.manual_seed(C + H + k + s)
    x = torch.randn(2, C, H, W, device=DEV, dtype=torch.bfloat16)
    xq, sx, _, _, _, (B, Ho, Wo) = ops().conv_act_quant(x, (k, k), (s, s), (p, p), (1, 1), "int8")
    cols = torch.nn.functional.unfold(x, kernel_size=k, padding=p, stride=s).transpose(1, 2).reshape(B * Ho * Wo, -1).contiguous()
    xq2, sx2, _, _, _ = ops().act_quant(cols, "int8")
    assert torch.equal(xq, xq2) and torch.equal(sx, sx2)
    # channels_last input: same logical tensor, different strides
    xq3, sx3, _, _, _, _ = ops().conv_act_quant(x.contiguous(memory_format=torch.channels_last), (k, k), (s, s), (p, p), (1, 1), "int8")
    assert torch.equal(xq, xq3) and torch.equal(sx, sx3)


def test_conv_unsupported_cases_fail_loudly():
    from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
    conv = torch.nn.Conv2d(64, 64, 3, padding=1, groups=2).to(torch.bfloat16)
    layer, _ = sdnq_quantize_layer(conv, SDNQConfig(weights_dtype="int8", quant_conv=True, use_quantized_matmul=True, use_quantized_matmul_conv=True))
    layer = layer.to(DEV)
    x = torch.randn(1, 64, 16, 16, device=DEV, dtype=torch.bfloat16)
    if layer.forward_func.__name__.endswith("_matmul"):
        with pytest.raises(NotImplementedError):
            layer(x)
