"""torch.library registration (sdnq_b200::*): schemas exist and the fake implementations describe the outputs, so the kernels can
be traced into torch.compile / export graphs without a GPU (the reference registers its Triton kernels as `sdnq::*` ops the same
way, kernels/triton_scaled_mm.py:239)."""
import pytest
import torch
from torch._subclasses import FakeTensorMode

import sdnq_b200.custom_ops  # noqa: F401  (registers the ops)


def test_ops_are_registered_with_schemas():
    for name in ("scaled_mm", "act_quant", "dequant_rowwise", "linear_small_m"):
        op = getattr(torch.ops.sdnq_b200, name)
        assert "Tensor" in str(op.default._schema)


def test_fake_implementations_describe_outputs():
    with FakeTensorMode():
        x = torch.empty(3, 40, 640, device="cuda", dtype=torch.bfloat16)
        xq, sx = torch.ops.sdnq_b200.act_quant(x, "int8", 0)
        assert xq.shape == (120, 640) and xq.dtype == torch.int8 and sx.shape == (120,) and sx.dtype == torch.float32 and xq.device.type == "cuda"
        wq = torch.empty(1280, 640, device="cuda", dtype=torch.int8)
        sw = torch.empty(1280, device="cuda")
        y = torch.ops.sdnq_b200.scaled_mm(xq, wq, sx, sw, None, torch.bfloat16)
        assert y.shape == (120, 1280) and y.dtype == torch.bfloat16
        xf, sf = torch.ops.sdnq_b200.act_quant(x, "float8_e4m3fn", 128)
        assert xf.dtype == torch.float8_e4m3fn
        w = torch.ops.sdnq_b200.dequant_rowwise(wq, "int8", sw, torch.bfloat16)
        assert w.shape == (1280, 640) and w.dtype == torch.bfloat16
        w_t = torch.ops.sdnq_b200.dequant_rowwise(wq.t(), "int8", sw, torch.float16)          # K-major [K,N] view of the same storage
        assert w_t.shape == (1280, 640)
        z = torch.ops.sdnq_b200.linear_small_m(x[:, :1].reshape(3, 640), wq, sw, None, None)
        assert z.shape == (3, 1280) and z.dtype == torch.bfloat16


def test_ops_trace_under_make_fx():
    from torch.fx.experimental.proxy_tensor import make_fx

    def f(x, wq, sw):
        xq, sx = torch.ops.sdnq_b200.act_quant(x, "int8", 0)
        return torch.ops.sdnq_b200.scaled_mm(xq, wq, sx, sw, None, torch.bfloat16)

    with FakeTensorMode() as mode:
        args = (torch.empty(64, 256, device="cuda", dtype=torch.bfloat16), torch.empty(128, 256, device="cuda", dtype=torch.int8),
                torch.empty(128, device="cuda"))
    gm = make_fx(f, tracing_mode="fake")(*args)
    targets = [str(n.target) for n in gm.graph.nodes if n.op == "call_function"]
    assert any("sdnq_b200.act_quant" in t for t in targets) and any("sdnq_b200.scaled_mm" in t for t in targets)


@pytest.mark.gpu
def test_custom_ops_run_the_kernels():
    from sdnq_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(64, 256, device="cuda", dtype=torch.bfloat16)
    wq = torch.randint(-128, 128, (128, 256), device="cuda", dtype=torch.int8)
    sw = torch.rand(128, device="cuda") * 0.01
    xq, sx = torch.ops.sdnq_b200.act_quant(x, "int8", 0)
    y = torch.ops.sdnq_b200.scaled_mm(xq, wq, sx, sw, None, torch.bfloat16)
    ref_q, ref_s, _, _, _ = ops.act_quant(x, "int8")
    assert torch.equal(xq, ref_q) and torch.equal(sx, ref_s)
    assert torch.equal(y, ops.scaled_mm(ref_q, wq, ref_s, sw, None, torch.bfloat16))
    w = torch.ops.sdnq_b200.dequant_rowwise(wq, "int8", sw, torch.bfloat16)
    assert torch.equal(w, ops.dequant(wq, "int8", sw, None, 128, 256, -1, torch.bfloat16))
    z = torch.ops.sdnq_b200.linear_small_m(x[:4], wq, sw, None, None)
    assert torch.equal(z, ops.linear_small_m(x[:4], wq, sw))
