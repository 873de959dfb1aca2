"""Run one kernel a few times (for ncu captures).  usage: python tools/run_one.py gemm M N K [int8|fp8] | actq M K [G mode] | dequant N K wdtype gs |
   dequant_svd N K | gemv M N K [fp8] | conv C H W N k"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import ops

kind = sys.argv[1]
dev = "cuda"
if kind == "gemm":
    M, N, K = map(int, sys.argv[2:5])
    fp8 = len(sys.argv) > 5 and sys.argv[5] == "fp8"
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    b = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
    if fp8:
        a = a.float().clamp(-8, 8).to(torch.float8_e4m3fn)
        b = b.float().clamp(-8, 8).to(torch.float8_e4m3fn)
    sx = torch.rand(M, device=dev) * 0.01
    sw = torch.rand(N, device=dev) * 0.01
    bias = torch.randn(N, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.scaled_mm(a, b, sx, sw, bias, torch.bfloat16)
elif kind == "actq":
    M, K = map(int, sys.argv[2:4])
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.act_quant(x, sys.argv[5] if len(sys.argv) > 5 else "int8", hadamard_group=int(sys.argv[4]) if len(sys.argv) > 4 else 0)
elif kind == "dequant":
    N, K = map(int, sys.argv[2:4])
    wd, gs = sys.argv[4], int(sys.argv[5])
    from sdnq_b200.common import dtype_dict
    bits = dtype_dict[wd]["num_bits"]
    w = torch.randint(0, 256, (N * K * bits // 8,), dtype=torch.uint8, device=dev)
    scale = torch.rand(N * (K // gs), device=dev) * 0.01
    zp = torch.rand(N * (K // gs), device=dev) if dtype_dict[wd]["is_unsigned"] else None
    for _ in range(4):
        ops.dequant(w, wd, scale, zp, N, K, gs, torch.bfloat16)
if kind == "dequant_svd":
    N, K = map(int, sys.argv[2:4])
    w = torch.randint(0, 256, (N * K // 2,), dtype=torch.uint8, device=dev)
    scale = torch.rand(N * (K // 128), device=dev) * 0.01
    up = torch.randn(N, 32, device=dev, dtype=torch.bfloat16) * 0.1
    down = (torch.randn(K, 32, device=dev, dtype=torch.bfloat16) * 0.1).t()
    for _ in range(4):
        ops.dequant(w, "int4", scale, None, N, K, 128, torch.bfloat16, svd_up=up, svd_down=down)
if kind == "gemv":
    M, N, K = map(int, sys.argv[2:5])
    q = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
    if len(sys.argv) > 5 and sys.argv[5] == "fp8":
        q = (q.float() / 4).to(torch.float8_e4m3fn)
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    sw = torch.rand(N, device=dev) * 0.01
    for _ in range(4):
        ops.linear_small_m(x, q, sw)
if kind == "conv":
    C, H, W, N, k = map(int, sys.argv[2:7])
    x = torch.randn(1, C, H, W, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.conv_act_quant(x, (k, k), (1, 1), (k // 2, k // 2), (1, 1), "int8")
torch.cuda.synchronize()
