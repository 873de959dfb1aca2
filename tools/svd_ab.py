"""int4 g128 + SVD r32 dequant (K3s) per SD-XL weight shape: tensor-core rank-r update vs the CUDA-core generic kernel.
Graph of `count` distinct weights (> L2), replay timed with CUDA events.   python tools/svd_ab.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import ops
from tools.shape_breakdown import graph_time

DEV = "cuda"
for (N, K, layers) in [(640, 640, 350), (640, 2048, 20), (5120, 640, 10), (640, 2560, 10), (1280, 1280, 252), (1280, 2048, 120), (10240, 1280, 60), (1280, 5120, 60)]:
    count = max(8, min(64, int(300e6 // (N * K * 2))))
    ws = [torch.randint(0, 256, (N * K // 2,), dtype=torch.uint8, device=DEV) for _ in range(count)]
    scale = torch.rand(N, K // 128, 1, device=DEV) * 0.01 + 1e-3
    up = (torch.randn(N, 32, device=DEV) * 0.1).to(torch.bfloat16)
    down = (torch.randn(K, 32, device=DEV) * 0.1).to(torch.bfloat16).t()
    res = {}
    for tag, env in (("tensor-core", "1"), ("cuda-core", "0")):
        os.environ["SDNQ_B200_SVD_TC"] = env
        outs = []

        def run():
            outs.clear()
            for w in ws:
                outs.append(ops.dequant(w, "int4", scale, None, N, K, 128, torch.bfloat16, svd_up=up, svd_down=down))
        res[tag] = graph_time(run) / count * 1e3
    os.environ.pop("SDNQ_B200_SVD_TC", None)
    by = N * K / 2 + scale.numel() * 4 + 2 * 32 * (N + K) + 2 * N * K
    print(f"dequant int4 g128 + svd32 {N:6d}x{K:5d} (x{layers:3d}/step): tensor-core {res['tensor-core']:7.2f} us ({by / res['tensor-core'] / 1e3:5.0f} GB/s)   "
          f"cuda-core {res['cuda-core']:7.2f} us ({by / res['cuda-core'] / 1e3:5.0f} GB/s)", flush=True)
