"""Sibling projections as one grouped launch (sdnq_b200/siblings.py): bit-identical to the layers' own forwards, and safe when the
guess about shared inputs is wrong."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"

CONFIGS = {
    "int8": dict(weights_dtype="int8", use_quantized_matmul=True),
    "uint8": dict(weights_dtype="uint8", use_quantized_matmul=True),
    "fp8_hadamard": dict(weights_dtype="float8_e4m3fn", use_quantized_matmul=True, use_hadamard=True, hadamard_group_size=128),
    "int4_rowwise": dict(weights_dtype="int4", use_quantized_matmul=True, group_size=-1),
    "uint4_g64": dict(weights_dtype="uint4", use_quantized_matmul=True, group_size=64),
    "int6": dict(weights_dtype="int6", use_quantized_matmul=True),
}


def make_layers(cfg, K, ns, bias=True, seed=0):
    from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
    torch.manual_seed(seed)
    out = []
    for n in ns:
        lin = torch.nn.Linear(K, n, bias=bias).to(torch.bfloat16)
        layer, _ = sdnq_quantize_layer(lin, SDNQConfig(**CONFIGS[cfg]))
        out.append(layer.to(DEV))
    return out


@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("M,K,ns", [(1024, 1280, (1280, 1280, 1280)), (77, 2048, (640, 640)), (300, 384, (256, 144, 528))])
def test_grouped_forward_is_bit_identical(cfg, M, K, ns, monkeypatch):
    from sdnq_b200 import group_siblings
    layers = make_layers(cfg, K, ns)
    x = torch.randn(2, M // 2, K, device=DEV, dtype=torch.bfloat16)
    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "0")
    ref = [layer(x.clone()) for layer in layers]
    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "1")
    group = group_siblings(layers)
    assert group is not None
    from sdnq_b200 import _lib
    layers[0](x.clone())                                 # builds the stacked operand (copies: not counted below)
    _lib.launch_count(reset=True)
    got = [layer(x) for layer in layers]
    launches = _lib.launch_count()
    assert launches == 2, launches                       # one K2 + one grouped K1 for all siblings
    for g, r in zip(got, ref):
        assert g.shape == r.shape and g.is_contiguous() and torch.equal(g, r)
    # a second round with a new tensor, called in another order
    x2 = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "0")
    ref2 = [layer(x2) for layer in layers]
    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "1")
    got2 = [layers[i](x2) for i in reversed(range(len(layers)))][::-1]
    for g, r in zip(got2, ref2):
        assert torch.equal(g, r)


def test_wrong_guess_about_shared_inputs_is_only_slower(monkeypatch):
    """to_q is fed the hidden states, to_k / to_v something else of the same width; inputs edited in place between calls."""
    from sdnq_b200 import group_siblings
    layers = make_layers("int8", 640, (640, 640, 640))
    xa = torch.randn(256, 640, device=DEV, dtype=torch.bfloat16)
    xb = torch.randn(256, 640, device=DEV, dtype=torch.bfloat16)
    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "0")
    ref = [layers[0](xa), layers[1](xb), layers[2](xb)]
    xa2 = xa * 2
    ref_edit = layers[1](xa2)
    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "1")
    assert group_siblings(layers) is not None
    got = [layers[0](xa), layers[1](xb), layers[2](xb)]
    for g, r in zip(got, ref):
        assert torch.equal(g, r)
    q = layers[0](xa)
    xa.mul_(2)                                           # same storage, new version: the pending k / v results must not be served
    assert torch.equal(layers[1](xa), ref_edit)
    assert torch.equal(q, ref[0])
    # small-M calls (rows < 32) bypass the group
    xs = torch.randn(4, 640, device=DEV, dtype=torch.bfloat16)
    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "0")
    r_small = layers[2](xs)
    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "1")
    assert torch.equal(layers[2](xs), r_small)


def test_group_follows_weight_replacement_and_graph_capture(monkeypatch):
    from sdnq_b200 import group_siblings
    layers = make_layers("int8", 1280, (1280, 1280))
    assert group_siblings(layers) is not None
    x = torch.randn(512, 1280, device=DEV, dtype=torch.bfloat16)
    y0 = [layer(x) for layer in layers]
    # replace one sibling's weight (what load_state_dict(assign=True) does): the shared operand is rebuilt
    other = make_layers("int8", 1280, (1280,), seed=7)[0]
    layers[1].weight = other.weight
    layers[1].scale = other.scale
    y1 = [layer(x) for layer in layers]
    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "0")
    alone = layers[1](x.clone())
    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "1")
    assert torch.equal(y1[0], y0[0]) and torch.equal(y1[1], alone) and not torch.equal(y1[1], y0[1])
    # captured: both launches are in the graph, replay reproduces the eager result for new input contents
    static_x = torch.randn(512, 1280, device=DEV, dtype=torch.bfloat16)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        [layer(static_x) for layer in layers]
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = [layer(static_x) for layer in layers]
    static_x.copy_(x)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(outs[0], y1[0]) and torch.equal(outs[1], y1[1])


def test_group_with_unused_results_dissolves_itself(monkeypatch):
    """a pooled group (to_k / to_v of two blocks) whose second block is never fed the same tensor stops computing for nothing."""
    from sdnq_b200 import group_siblings
    layers = make_layers("int8", 256, (256, 256, 256, 256))
    group = group_siblings(layers)
    assert group is not None
    for i in range(8):
        x = torch.randn(64, 256, device=DEV, dtype=torch.bfloat16)
        monkeypatch.setenv("SDNQ_B200_SIBLINGS", "0")
        ref = [layers[0](x), layers[1](x)]
        monkeypatch.setenv("SDNQ_B200_SIBLINGS", "1")
        assert torch.equal(layers[0](x), ref[0]) and torch.equal(layers[1](x), ref[1])
    assert group.dead and all("_sdnq_siblings" not in layer.__dict__ for layer in layers)


def test_dequant_path_siblings_share_one_batched_gemm(monkeypatch):
    """to_q / to_k / to_v on the dequant path: once the batched dequantiser has put their weights next to each other, their three
    F.linear calls become one strided-batched library GEMM; outputs equal the members' own F.linear up to the library's accumulation
    order (different cuBLAS kernel: one bf16 ulp)."""
    from sdnq_b200 import SDNQConfig, group_siblings, prefetch, sdnq_quantize_layer
    from sdnq_b200.siblings import DequantSiblingGroup
    torch.manual_seed(5)
    cfg = SDNQConfig(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32, svd_steps=2)
    qkv = [sdnq_quantize_layer(torch.nn.Linear(640, 640, bias=(i != 1)).to(torch.bfloat16), cfg)[0].to(DEV) for i in range(3)]
    out_proj = sdnq_quantize_layer(torch.nn.Linear(640, 640).to(torch.bfloat16), cfg)[0].to(DEV)
    x = torch.randn(2, 128, 640, device=DEV, dtype=torch.bfloat16)

    def block(inp):
        q, k, v = (layer(inp) for layer in qkv)
        return out_proj(q + k + v), (q, k, v)

    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "0")
    prefetch.reset()
    ref, ref_qkv = block(x)
    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "1")
    group = group_siblings(qkv)
    assert type(group) is DequantSiblingGroup
    calls = {"n": 0}
    real_bmm, real_baddbmm = torch.bmm, torch.baddbmm
    monkeypatch.setattr(torch, "bmm", lambda *a, **k: (calls.__setitem__("n", calls["n"] + 1), real_bmm(*a, **k))[1])
    monkeypatch.setattr(torch, "baddbmm", lambda *a, **k: (calls.__setitem__("n", calls["n"] + 1), real_baddbmm(*a, **k))[1])
    prefetch.reset()
    for step in range(4):
        calls["n"] = 0
        got, got_qkv = block(x)
        for g, r in zip(got_qkv + (got,), ref_qkv + (ref,)):
            assert g.shape == r.shape and g.is_contiguous()
            assert float((g.float() - r.float()).abs().max()) <= 2.0 ** -6 * float(r.float().abs().max()), step
        if step >= 2:
            assert calls["n"] == 1, (step, calls["n"])           # the three projections ran as one batched GEMM
    # a member fed something else still gets the right answer
    other = torch.randn(2, 128, 640, device=DEV, dtype=torch.bfloat16)
    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "0")
    want = qkv[1](other)
    monkeypatch.setenv("SDNQ_B200_SIBLINGS", "1")
    qkv[0](x)
    assert float((qkv[1](other).float() - want.float()).abs().max()) <= 2.0 ** -6 * float(want.float().abs().max())
