import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.bench_kernels import timeit
from sdnq_b200 import ops
dev = "cuda"
for mb in (8, 32, 100, 400):
    n = mb * 1024 * 1024 // 2
    a = torch.randn(n, device=dev, dtype=torch.bfloat16); b = torch.empty_like(a)
    med, best = timeit(lambda: b.copy_(a))
    med2, best2 = timeit(lambda: b.copy_(a), flush=False)
    print(f"copy {mb} MB: flush med {2*mb/1024/med*1e3:.0f} GB/s best {2*mb/1024/best*1e3:.0f} | noflush med {2*mb/1024/med2*1e3:.0f} GB/s  ({med*1e3:.1f} us / {med2*1e3:.1f} us)")
for (M, K) in [(1024, 1280), (4096, 640), (16384, 3072), (16384, 12288)]:
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    med, _ = timeit(lambda: ops.act_quant(x, "int8"))
    med2, _ = timeit(lambda: ops.act_quant(x, "int8"), flush=False)
    print(f"act_quant {M}x{K}: flush {3*M*K/med/1e6:.0f} GB/s ({med*1e3:.1f} us) | noflush {3*M*K/med2/1e6:.0f} GB/s ({med2*1e3:.1f} us)")
for (N, K) in [(10240, 1280), (12288, 3072)]:
    w = torch.randint(0, 256, (N * K // 2,), dtype=torch.uint8, device=dev)
    scale = torch.rand(N * (K // 128), device=dev) * 0.01
    med, _ = timeit(lambda: ops.dequant(w, "int4", scale, None, N, K, 128, torch.bfloat16))
    med2, _ = timeit(lambda: ops.dequant(w, "int4", scale, None, N, K, 128, torch.bfloat16), flush=False)
    by = N * K / 2 + 4 * N * K / 128 + 2 * N * K
    print(f"dequant int4 {N}x{K}: flush {by/med/1e6:.0f} GB/s ({med*1e3:.1f} us) | noflush {by/med2/1e6:.0f} GB/s ({med2*1e3:.1f} us)")

print("--- fp8 + hadamard act_quant")
for (M, K) in [(16384, 3072), (18432, 15360)]:
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    med, _ = timeit(lambda: ops.act_quant(x, "float8_e4m3fn", hadamard_group=256))
    print(f"act_quant fp8+had {M}x{K}: {3*M*K/med/1e6:.0f} GB/s ({med*1e3:.1f} us)")
