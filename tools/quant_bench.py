"""K8 (load-time quantise + pack) against the eager tensor ops it replaces, FLUX / SD-XL weight shapes.   python tools/quant_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import ops, packing, quant_math

DEV = "cuda"


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    t1.synchronize()
    return t0.elapsed_time(t1) / reps * 1e3


for (N, K) in [(3072, 3072), (12288, 3072), (3072, 12288), (1280, 1280), (10240, 1280)]:
    for wd, gs in (("int8", -1), ("uint4", 32), ("int4", 128), ("int6", -1), ("int3", 64)):
        w = torch.randn(N, K, device=DEV, dtype=torch.bfloat16)
        info = quant_math.dtype_dict[wd]

        def eager():
            view = w.float().view(N, 1 if gs <= 0 else K // gs, -1)
            q, s, z = quant_math.quantize_weight(view, -1, wd)
            return packing.pack_int(q, wd) if info["is_packed"] else q

        t_k = timed(lambda: ops.quantize_weight(w, wd, gs))
        t_e = timed(eager, reps=3)
        by = N * K * (2 + info["num_bits"] / 8)
        print(f"{wd:>6s} g={gs:4d} {N:6d}x{K:6d}: kernel {t_k:8.1f} us ({by / t_k / 1e6:5.2f} TB/s of bf16 in + packed out)   eager ops {t_e:9.1f} us   x{t_e / t_k:5.1f}", flush=True)
