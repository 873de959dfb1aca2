"""K2 (activation pre-pass) timing: rotating inputs larger than L2, back-to-back launches between two CUDA events.
    python tools/actq_ab.py            (SDNQ_B200_HADAMARD_BUTTERFLY=1 selects the shuffle-butterfly rotation for A/B)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import ops

DEV = "cuda"
tag = "butterfly" if os.environ.get("SDNQ_B200_HADAMARD_BUTTERFLY", "0") not in ("", "0") else "tensor-core"
for (M, K) in [(16384, 3072), (18432, 15360), (16384, 12288), (2048, 3072), (4096, 640), (1024, 1280)]:
    nbuf = max(2, int(600e6 // (M * K * 2)) + 1)
    xs = [torch.randn(M, K, device=DEV, dtype=torch.bfloat16) for _ in range(min(nbuf, 8))]
    for mode, hg in [("int8", 0), ("float8_e4m3fn", 0), ("float8_e4m3fn", 256 if K % 256 == 0 else 128), ("int8", 128), ("int8", 64), ("uint8", 256 if K % 256 == 0 else 128)]:
        for x in xs[:2]:
            ops.act_quant(x, mode, hadamard_group=hg)
        torch.cuda.synchronize()
        iters = 5 * len(xs)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(iters):
            ops.act_quant(xs[i % len(xs)], mode, hadamard_group=hg)
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) * 1000 / iters
        print(f"[{tag}] act_quant M={M} K={K} {mode} hadamard={hg}: {us:.1f} us  {3.0 * M * K / us / 1e3:.0f} GB/s", flush=True)
