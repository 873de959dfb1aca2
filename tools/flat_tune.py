"""dequant_flat_kernel tuning knobs (CTAs per SM, streaming stores) on large int4 / int8 weights.  python tools/flat_tune.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import ops
from tools.shape_breakdown import graph_time

DEV = "cuda"
for (N, K, wd, gs) in [(12288, 3072, "int4", 128), (18432, 3072, "int4", 128), (12288, 3072, "int8", -1), (10240, 1280, "int4", 128)]:
    bits = 4 if wd == "int4" else 8
    count = max(4, int(800e6 // (N * K * 2)))
    ws = [torch.randint(0, 256, (N * K * bits // 8,), dtype=torch.uint8, device=DEV) for _ in range(count)]
    if wd == "int8":
        ws = [w.view(torch.int8) for w in ws]
    groups = K // gs if gs > 0 else 1
    scale = torch.rand((N, groups, 1) if groups > 1 else (N, 1), device=DEV) * 0.01 + 1e-3
    by = N * K * bits // 8 + scale.numel() * 4 + 2 * N * K
    line = f"{wd} g{gs} {N}x{K}:"
    for grid in (32, 64, 128, 4096):
        for streaming in (0, 1):
            os.environ["SDNQ_B200_FLAT_GRID"] = str(grid)
            os.environ["SDNQ_B200_FLAT_STREAMING"] = str(streaming)
            outs = []

            def run():
                outs.clear()
                for w in ws:
                    outs.append(ops.dequant(w, wd, scale, None, N, K, gs, torch.bfloat16))
            ms = graph_time(run) / count
            line += f"  g{grid}{'s' if streaming else ' '} {by / ms / 1e6:5.0f}"
    print(line + "  GB/s", flush=True)
