"""grid cap of the row-walking dequant kernels (rotated 8-bit, packed sub-byte, generic).  python tools/dq_grid_tune.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import ops
from tools.shape_breakdown import graph_time

DEV = "cuda"
for (N, K, wd, gs, hg) in [(18432, 3072, "float8_e4m3fn", -1, 256), (12288, 3072, "int8", -1, 128), (12288, 3072, "uint3", 64, 0), (12288, 3072, "int5", 128, 0),
                           (12288, 3072, "float6_e3m2fn", -1, 0), (12288, 3072, "float8_e4m3fn", -1, 0)]:
    bits = 8 if "8" in wd.split("_")[0] else int("".join(ch for ch in wd.split("_")[0] if ch.isdigit()))
    count = max(4, int(600e6 // (N * K * 2)))
    ws = [torch.randint(0, 256, (N * K * bits // 8,), dtype=torch.uint8, device=DEV) for _ in range(count)]
    if wd == "int8":
        ws = [w.view(torch.int8) for w in ws]
    if wd == "float8_e4m3fn":
        ws = [(w & 0x77).view(torch.float8_e4m3fn) for w in ws]
    groups = K // gs if gs > 0 else 1
    scale = torch.rand((N, groups, 1) if groups > 1 else (N, 1), device=DEV) * 0.01 + 1e-3
    zp = torch.randn_like(scale) * 0.01 if wd.startswith("u") else None
    by = N * K * bits // 8 + scale.numel() * 4 * (2 if zp is not None else 1) + 2 * N * K
    line = f"{wd} g{gs} had{hg} {N}x{K}:"
    for grid in (8, 16, 32, 64, 256):
        os.environ["SDNQ_B200_DQ_GRID"] = str(grid)
        outs = []

        def run():
            outs.clear()
            for w in ws:
                outs.append(ops.dequant(w, wd, scale, zp, N, K, gs, torch.bfloat16, hadamard_group=hg))
        ms = graph_time(run) / count
        line += f"  g{grid} {by / ms / 1e6:5.0f}"
    print(line + "  GB/s", flush=True)
