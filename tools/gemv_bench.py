"""K5 (small-M Linear) timing against the dequantise + bf16 GEMM path it replaces.  python tools/gemv_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import ops
from tools.shape_breakdown import graph_time

DEV = "cuda"
for (M, N, K, fp8) in [(4, 18432, 3072, True), (4, 9216, 3072, True), (1, 1280, 1280, False), (1, 1280, 2816, False), (2, 320, 1280, False), (16, 18432, 3072, True), (31, 4096, 4096, False)]:
    count = max(4, min(64, int(400e6 // (N * K))))
    x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=DEV) for _ in range(count)]
    if fp8:
        ws = [(w.float() / 4).to(torch.float8_e4m3fn) for w in ws]
    sw = torch.rand(N, device=DEV) * 0.01 + 1e-3
    bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)
    wd = "float8_e4m3fn" if fp8 else "int8"

    def gemv():
        for w in ws:
            ops.linear_small_m(x, w, sw, bias=bias)

    def deq():
        for w in ws:
            W = ops.dequant(w, wd, sw.view(N, 1), None, N, K, -1, torch.bfloat16)
            torch.nn.functional.linear(x, W, bias)
    t1 = graph_time(gemv) / count * 1e3
    t2 = graph_time(deq) / count * 1e3
    print(f"M={M:2d} N={N:6d} K={K:5d} {wd:14s}: K5 {t1:7.2f} us ({N * K / t1 / 1e3:6.0f} GB/s of codes)   dequant + bf16 GEMM {t2:7.2f} us", flush=True)

# ---- K5p (packed / group-wise weights) against dequantise + bf16 GEMM; first hardware numbers are a round-2 item
for (M, N, K, wd, bits, gs) in [(4, 18432, 3072, "int4", 4, 128), (4, 18432, 3072, "uint4", 4, 64), (1, 1280, 1280, "int4", 4, 128), (2, 1280, 2816, "int4", 4, 128),
                                (4, 9216, 3072, "int2", 2, 16), (16, 18432, 3072, "int4", 4, 128)]:
    count = max(4, min(64, int(400e6 // (N * K))))
    x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    ws = [torch.randint(0, 256, (N * K * bits // 8,), dtype=torch.uint8, device=DEV) for _ in range(count)]
    scale = torch.rand(N, K // gs, 1, device=DEV) * 0.01 + 1e-3
    zp = torch.randn(N, K // gs, 1, device=DEV) * 0.01 if wd.startswith("u") else None
    bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)

    def gemvp():
        for w in ws:
            ops.linear_small_m_packed(x, w, wd, scale, zp, N, K, bias=bias)

    def deqp():
        for w in ws:
            W = ops.dequant(w, wd, scale, zp, N, K, gs, torch.bfloat16)
            torch.nn.functional.linear(x, W, bias)
    t1 = graph_time(gemvp) / count * 1e3
    t2 = graph_time(deqp) / count * 1e3
    print(f"M={M:2d} N={N:6d} K={K:5d} {wd:6s} g{gs:<4d}: K5p {t1:7.2f} us ({N * K * bits / 8 / t1 / 1e3:6.0f} GB/s of stored bytes)   dequant + bf16 GEMM {t2:7.2f} us", flush=True)
