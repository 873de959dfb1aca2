"""K1 int8: whole-tile scheduling vs stream-K (SDNQ_B200_STREAMK=0 / 1), per shape; `count` weights (> L2) in one CUDA graph.   python tools/streamk_ab.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import ops
from tools.shape_breakdown import graph_time

DEV = "cuda"
for (M, N, K) in [(1024, 1280, 5120), (4096, 640, 2560), (1024, 1280, 1280), (4096, 640, 640), (1024, 1280, 2560), (1024, 640, 5120), (2048, 1280, 5120),
                  (1024, 2560, 5120), (512, 1280, 5120), (1024, 1280, 10240), (256, 1280, 5120)]:
    count = max(4, min(64, int(400e6 // (N * K))))
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=DEV)
    ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=DEV) for _ in range(count)]
    sx, sw = torch.rand(M, device=DEV) * 0.01, torch.rand(N, device=DEV) * 0.01
    bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)
    res = {}
    for mode in ("0", "1"):
        os.environ["SDNQ_B200_STREAMK"] = mode

        def run():
            for w in ws:
                ops.scaled_mm(a, w, sx, sw, bias, torch.bfloat16)
        res[mode] = graph_time(run) / count * 1e3
    os.environ.pop("SDNQ_B200_STREAMK", None)
    tiles, kb = ((M + 127) // 128) * ((N + 127) // 128), (K + 127) // 128
    fl = 2 * M * N * K / 1e6
    print(f"{M:5d}x{N:5d}x{K:5d}  tiles {tiles:4d} x {kb:3d} k-blocks: whole tiles {res['0']:7.2f} us ({fl / res['0']:7.1f} TF)   stream-K {res['1']:7.2f} us ({fl / res['1']:7.1f} TF)", flush=True)
