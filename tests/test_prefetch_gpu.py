"""Batched dequantisation ahead of the GEMMs (sdnq_b200/prefetch.py, sdnq_b200_dequant_batch_*): identical bits to the per-layer
dequant path, whatever the call order does."""
import numpy as np
import pytest
import torch

from oracle import sdnq_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_batched_dequant_kernel_equals_per_layer_kernel():
    from sdnq_b200 import ops
    rng = np.random.default_rng(5)
    jobs, ref = [], []
    for (N, K, rank, gs, wd) in [(128, 256, 32, 128, "int4"), (300, 640, 32, 128, "uint4"), (520, 1280, 16, 8, "int4"), (257, 2048, 64, 64, "uint4"),
                                 (1280, 1280, 32, -1, "int4"), (640, 640, 32, 128, "int4"), (8, 32, 16, 32, "uint4"), (5120, 640, 32, 128, "int4")]:
        groups = 1 if gs <= 0 else K // gs
        packed = torch.from_numpy(O.pack_uint(rng.integers(0, 16, size=(N, K)), 4).astype(np.uint8)).to(DEV)
        scale = torch.from_numpy((rng.random((N, groups, 1)) * 0.02 + 1e-3).astype(np.float32)).to(DEV)
        zp = torch.from_numpy(rng.standard_normal((N, groups, 1)).astype(np.float32) * 0.1).to(DEV) if wd == "uint4" else None
        up = torch.from_numpy(rng.standard_normal((N, rank)).astype(np.float32) * 0.1).to(torch.bfloat16).to(DEV)
        down = torch.from_numpy(rng.standard_normal((K, rank)).astype(np.float32) * 0.1).to(torch.bfloat16).to(DEV).t()      # [r,K] stored K-major
        jobs.append(dict(weight=packed, weights_dtype=wd, scale=scale, zero_point=zp, N=N, K=K, group_size=gs, svd_up=up, svd_down=down, svd_layout_matmul=False))
        ref.append(ops.dequant(packed, wd, scale, zp, N, K, gs, torch.bfloat16, svd_up=up, svd_down=down))
    offs = ops.dequant_batch_bytes([(j["N"], j["K"]) for j in jobs])
    slab = torch.empty(offs[-1], dtype=torch.uint8, device=DEV)
    import os
    for tn in ("64", "128", "256"):
        os.environ["SDNQ_B200_SVD_BATCH_TN"] = tn
        try:
            plan = ops.dequant_batch_plan(jobs, slab)
        finally:
            os.environ.pop("SDNQ_B200_SVD_BATCH_TN")
        slab.fill_(0xFF)
        ops.dequant_batch_run(plan)
        torch.cuda.synchronize()
        for got, want in zip(plan.outs, ref):
            assert torch.equal(got, want), tn
    # a weight the kernel does not cover is refused at plan time
    bad = dict(jobs[0], svd_up=jobs[0]["svd_up"].float())
    from sdnq_b200 import _lib
    with pytest.raises(_lib.SDNQKernelError):
        ops.dequant_batch_plan([bad], slab)


def make_stack(seed=0):
    from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
    torch.manual_seed(seed)
    layers = []
    for (K, N) in [(640, 640), (640, 1280), (1280, 1280), (1280, 5120), (5120, 640), (640, 256), (256, 640)]:
        lin = torch.nn.Linear(K, N).to(torch.bfloat16)
        layer, _ = sdnq_quantize_layer(lin, SDNQConfig(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32))
        layers.append(layer.to(DEV))
    return layers


def run_chain(layers, x, order=None):
    outs = {}
    h = x
    for i in (order or range(len(layers))):
        layer = layers[i]
        inp = h if h.shape[-1] == layer.sdnq_dequantizer.original_shape[1] else torch.ones(x.shape[0], layer.sdnq_dequantizer.original_shape[1], device=DEV, dtype=torch.bfloat16)
        h = layer(inp)
        outs[i] = h
    return outs


def test_prefetched_forward_is_bit_identical_and_batches_launches(monkeypatch):
    from sdnq_b200 import _lib, prefetch
    layers = make_stack()
    x = torch.randn(64, 640, device=DEV, dtype=torch.bfloat16)
    monkeypatch.setenv("SDNQ_B200_DEQUANT_PREFETCH", "0")
    ref = run_chain(layers, x)
    _lib.launch_count(reset=True)
    run_chain(layers, x)
    per_layer_launches = _lib.launch_count()
    assert per_layer_launches == len(layers)
    monkeypatch.setenv("SDNQ_B200_DEQUANT_PREFETCH", "1")
    prefetch.reset()
    for step in range(4):                                 # step 0 learns the order, step 1 plans, steps 2+ reuse the plans
        _lib.launch_count(reset=True)
        got = run_chain(layers, x)
        launches = _lib.launch_count()
        for i in ref:
            assert torch.equal(got[i], ref[i]), (step, i)
        if step >= 1:
            assert launches < per_layer_launches, (step, launches)
    # the order changes: still the right weights
    order = [3, 0, 6, 2, 5, 1, 4]
    monkeypatch.setenv("SDNQ_B200_DEQUANT_PREFETCH", "0")
    ref2 = run_chain(layers, x, order)
    monkeypatch.setenv("SDNQ_B200_DEQUANT_PREFETCH", "1")
    for step in range(3):
        got2 = run_chain(layers, x, order)
        for i in ref2:
            assert torch.equal(got2[i], ref2[i]), (step, i)
    # small budget: several batches per step, slabs recycled
    monkeypatch.setenv("SDNQ_B200_PREFETCH_MB", "16")
    prefetch.reset()
    for step in range(4):
        got = run_chain(layers, x)
        for i in ref:
            assert torch.equal(got[i], ref[i]), ("small budget", step, i)


def test_prefetch_follows_weight_replacement_and_graph_capture(monkeypatch):
    from sdnq_b200 import prefetch
    layers = make_stack(seed=3)
    x = torch.randn(128, 640, device=DEV, dtype=torch.bfloat16)
    prefetch.reset()
    for _ in range(3):
        run_chain(layers, x)
    other = make_stack(seed=9)[2]
    layers[2].weight, layers[2].scale, layers[2].svd_up, layers[2].svd_down = other.weight, other.scale, other.svd_up, other.svd_down
    monkeypatch.setenv("SDNQ_B200_DEQUANT_PREFETCH", "0")
    ref = run_chain(layers, x)
    monkeypatch.setenv("SDNQ_B200_DEQUANT_PREFETCH", "1")
    for step in range(3):
        got = run_chain(layers, x)
        for i in ref:
            assert torch.equal(got[i], ref[i]), (step, i)
    # whole step captured: replay reproduces the eager result for new input contents
    static_x = torch.randn(128, 640, device=DEV, dtype=torch.bfloat16)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run_chain(layers, static_x)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = run_chain(layers, static_x)
    static_x.copy_(x)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    for i in ref:
        assert torch.equal(outs[i], ref[i]), i
    # and eager again afterwards
    got = run_chain(layers, x)
    for i in ref:
        assert torch.equal(got[i], ref[i]), i
