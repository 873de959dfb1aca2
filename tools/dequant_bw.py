"""HBM throughput of the dequant-only kernels inside a CUDA graph (GPU box): `count` distinct weights (footprint >> L2) are
dequantised back to back, the replay is timed with CUDA events.  Algorithmic bytes = stored codes + scales (+ zero points) read
+ bf16 weight written.  Not part of the product."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from sdnq_b200 import ops
from tools.shape_breakdown import graph_time

DEV = "cuda"


def main():
    # (weights_dtype, group_size, hadamard group for the un-rotate: 0 = none)
    cases = [("int4", 128, 0), ("uint4", 128, 0), ("int8", -1, 0), ("int4", 32, 0), ("uint3", 64, 0), ("int5", 128, 0), ("float8_e4m3fn", -1, 0),
             ("float6_e3m2fn", -1, 0), ("float8_e4m3fn", -1, 256), ("int8", -1, 128)]
    if "--rot" in sys.argv:
        cases = [c for c in cases if c[2]]
    for (N, K) in [(10240, 1280), (12288, 3072), (18432, 3072), (3072, 12288), (1280, 5120)]:
        for wd, gs, hg in cases:
            bits = 8 if "8" in wd else int("".join(ch for ch in wd.split("_")[0] if ch.isdigit()))
            count = max(4, int(600e6 // (N * K * 2)))
            nbytes = N * K * bits // 8
            ws = [torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device=DEV) for _ in range(count)]
            if wd == "int8":
                ws = [w.view(torch.int8) for w in ws]
            if wd == "float8_e4m3fn":
                ws = [(w & 0x77).view(torch.float8_e4m3fn) for w in ws]
            groups = K // gs if gs > 0 else 1
            scale = torch.rand((N, groups, 1) if groups > 1 else (N, 1), device=DEV) * 0.01 + 1e-3
            zp = torch.randn_like(scale) * 0.01 if wd.startswith("u") else None
            outs = []

            def run():
                outs.clear()
                for w in ws:
                    outs.append(ops.dequant(w, wd, scale, zp, N, K, gs, torch.bfloat16, hadamard_group=hg))

            ms = graph_time(run) / count
            by = nbytes + scale.numel() * 4 * (2 if zp is not None else 1) + 2 * N * K
            print(f"dequant {wd:14s} g{gs:4d} had{hg:3d} {N:6d}x{K:6d}: {ms * 1e3:8.2f} us  {by / ms / 1e6:7.0f} GB/s", flush=True)
            del ws, outs
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
