"""Run one kernel a few times (for ncu captures).  usage: python tools/run_one.py gemm M N K [int8|fp8] | actq M K [G mode] | dequant N K wdtype gs |
   dequant_svd N K | gemv M N K [fp8] | conv C H W N k | grouped M K N1 N2 ... [fp8] | svd_mm M N K | packed M N K wdtype | quantw N K wdtype gs |
   dequant_batch [MB] | svd_low M K r"""
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
if kind == "grouped":
    fp8 = sys.argv[-1] == "fp8"
    nums = [int(v) for v in sys.argv[2:] if v.isdigit()]
    M, K, ns = nums[0], nums[1], nums[2:]
    align = 256 if all(n % 256 == 0 for n in ns) else 128
    starts = [0]
    for n in ns:
        starts.append(starts[-1] + (n + align - 1) // align * align)
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    b = torch.randint(-128, 128, (starts[-1], K), dtype=torch.int8, device=dev)
    if fp8:
        a, b = a.float().clamp(-8, 8).to(torch.float8_e4m3fn), b.float().clamp(-8, 8).to(torch.float8_e4m3fn)
    sx, sw = torch.rand(M, device=dev) * 0.01, torch.rand(starts[-1], device=dev) * 0.01
    bias = torch.randn(starts[-1], device=dev)
    for _ in range(4):
        ops.scaled_mm_grouped(a, b, sx, sw, starts, ns, bias, torch.bfloat16)
if kind == "svd_mm":
    M, N, K = map(int, sys.argv[2:5])
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    b = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    down = torch.randn(32, K, device=dev, dtype=torch.bfloat16) * 0.05
    up = torch.randn(N, 32, device=dev, dtype=torch.bfloat16) * 0.05
    sx, sw = torch.rand(M, device=dev) * 0.01, torch.rand(N, device=dev) * 0.01
    bias = torch.randn(N, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        low = ops.svd_low(x, down)
        ops.scaled_mm_svd(a, b, sx, sw, low, up, bias, torch.bfloat16)
if kind == "svd_low":
    M, K, r = map(int, sys.argv[2:5])
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    down = torch.randn(r, K, device=dev, dtype=torch.bfloat16) * 0.05
    for _ in range(4):
        ops.svd_low(x, down)
if kind == "packed":
    M, N, K = map(int, sys.argv[2:5])
    wd = sys.argv[5]
    from sdnq_b200.common import dtype_dict
    info = dtype_dict[wd]
    packed = torch.randint(0, 256, (N * K * info["num_bits"] // 8,), dtype=torch.uint8, device=dev)
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    if not info["is_integer"]:
        a = a.float().clamp(-8, 8).to(torch.float8_e4m3fn)
    sx, sw = torch.rand(M, device=dev) * 0.01, torch.rand(N, device=dev) * 0.01
    kw = dict(zp=torch.rand(N, device=dev), rowsum=torch.zeros(M, dtype=torch.int32, device=dev)) if info["is_integer"] and info["is_unsigned"] else {}
    for _ in range(4):
        ops.scaled_mm_packed(a, packed, wd, N, sx, sw, None, torch.bfloat16, **kw)
if kind == "quantw":
    N, K = map(int, sys.argv[2:4])
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.quantize_weight(w, sys.argv[4], int(sys.argv[5]))
if kind == "dequant_batch":
    budget = (int(sys.argv[2]) if len(sys.argv) > 2 else 160) << 20
    shapes = [(1280, 1280)] * 8 + [(10240, 1280), (1280, 5120), (640, 640), (640, 640), (5120, 640), (640, 2560), (1280, 2048), (1280, 2048)]
    jobs, total = [], 0
    while True:
        N, K = shapes[len(jobs) % len(shapes)]
        if total + N * K * 2 > budget:
            break
        total += (N * K * 2 + 255) // 256 * 256
        jobs.append(dict(weight=torch.randint(0, 256, (N * K // 2,), dtype=torch.uint8, device=dev), weights_dtype="int4",
                         scale=torch.rand(N, K // 128, 1, device=dev) * 0.01, zero_point=None, N=N, K=K, group_size=128,
                         svd_up=torch.randn(N, 32, device=dev, dtype=torch.bfloat16) * 0.1, svd_down=(torch.randn(K, 32, device=dev, dtype=torch.bfloat16) * 0.1).t(),
                         svd_layout_matmul=False))
    slab = torch.empty(ops.dequant_batch_bytes([(j["N"], j["K"]) for j in jobs])[-1], dtype=torch.uint8, device=dev)
    plan = ops.dequant_batch_plan(jobs, slab)
    print(len(jobs), "weights", total >> 20, "MB of bf16 out")
    for _ in range(4):
        ops.dequant_batch_run(plan)
torch.cuda.synchronize()
