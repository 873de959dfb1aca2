import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import sdnq_oracle as O
from sdnq_b200 import ops
from tools.bench_kernels import timeit
dev = "cuda"
def run(N, K, rank, gs, wd="int4", bits=4, check=True):
    rng = np.random.default_rng(0)
    codes = rng.integers(0, 2 ** bits, size=(N, K))
    packed = torch.from_numpy(O.pack_uint(codes, bits).astype(np.uint8)).to(dev)
    groups = K // gs
    scale = torch.from_numpy((rng.random((N, groups, 1) if groups > 1 else (N, 1)) * 0.02 + 0.001).astype(np.float32)).to(dev)
    up = (torch.from_numpy(rng.standard_normal((N, rank)).astype(np.float32)) * 0.1).to(torch.bfloat16).to(dev)
    down_phys = (torch.from_numpy(rng.standard_normal((K, rank)).astype(np.float32)) * 0.1).to(torch.bfloat16).to(dev)
    f = lambda: ops.dequant(packed, wd, scale, None, N, K, gs if groups > 1 else -1, torch.bfloat16, svd_up=up, svd_down=down_phys.t())
    t0 = time.time(); W = f(); torch.cuda.synchronize(); print(f"{N}x{K} r{rank} g{gs}: first call {1e3*(time.time()-t0):.1f} ms")
    med, best = timeit(f, iters=5)
    by = N*K*bits/8 + 4*N*groups + 2*rank*(N+K) + 2*N*K
    print(f"   {med*1e3:.1f} us  {by/med/1e6:.0f} GB/s")
    if check:
        layer = O.Layer(packed.cpu().numpy(), scale.cpu().numpy(), None, up.float().cpu().numpy(), down_phys.t().float().cpu().numpy(), weights_dtype=wd,
                        quantized_weight_shape=[N, groups, gs] if groups > 1 else [N, K], result_shape=[N, K] if groups > 1 else None, group_size=gs if groups > 1 else -1)
        ref = O.dequantize(layer, dtype="bfloat16")
        got = W.float().cpu().numpy()
        err = np.abs(got - ref); bad = np.argwhere(err > 2.0**-8 * np.abs(ref).max(axis=-1, keepdims=True))
        print("   max err", err.max(), "bad", len(bad), bad[:5].tolist(), [ (got[tuple(b)], ref[tuple(b)]) for b in bad[:3]])
run(128, 256, 32, 128)
run(1280, 1280, 32, 1280)
run(1280, 1280, 32, 128)
run(10240, 1280, 32, 128, check=False)
run(1280, 5120, 32, 128, check=False)

# a real quantised layer through the public API
import copy
from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
lin = torch.nn.Linear(1280, 10240, device=dev, dtype=torch.bfloat16)
layer, _ = sdnq_quantize_layer(lin, SDNQConfig(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32))
d = layer.sdnq_dequantizer
print(layer.svd_up.shape, layer.svd_up.stride(), layer.svd_down.shape, layer.svd_down.stride(), layer.svd_up.dtype, layer.scale.shape)
from sdnq_b200 import _lib
_lib.launch_count(reset=True)
f = lambda: d(layer.weight, layer.scale, layer.zero_point, layer.svd_up, layer.svd_down)
f(); torch.cuda.synchronize(); print("launches", _lib.launch_count())
med, best = timeit(f, iters=5); print(f"layer dequant 10240x1280: {med*1e3:.1f} us")
x = torch.randn(1024, 1280, device=dev, dtype=torch.bfloat16)
med, best = timeit(lambda: layer(x), iters=5); print(f"layer forward M=1024: {med*1e3:.1f} us")
