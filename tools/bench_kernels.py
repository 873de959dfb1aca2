"""Per-kernel timing on the GPU box (CUDA events, L2 flushed between iterations).

    python tools/bench_kernels.py [--quick]

Prints one line per (kernel, shape): ours vs the cuBLASLt call the reference's eager path would make
(torch._int_mm / torch._scaled_mm), plus the measured cuBLASLt int8 / fp8 peak at 8192^3 that bench.py uses as the
tensor roofline denominator.  Not part of the product; numbers land in gpurun_out/kernels.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from sdnq_b200 import ops

DEV = "cuda"
FLUSH = None


def flush_l2():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    FLUSH.zero_()


def timeit(fn, iters=20, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    quick = "--quick" in sys.argv
    res = {}
    shapes = [(4096, 640, 640), (4096, 5120, 640), (4096, 640, 2560), (1024, 1280, 1280), (1024, 10240, 1280), (1024, 1280, 5120),
              (77, 1280, 2048), (16384, 3072, 3072), (16384, 12288, 3072), (16384, 3072, 12288), (18432, 3072, 15360), (8192, 8192, 8192)]
    if quick:
        shapes = shapes[:2] + shapes[7:8]
    for (M, N, K) in shapes:
        a8 = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=DEV)
        b8 = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=DEV)
        sx = torch.rand(M, device=DEV) * 0.01
        sw = torch.rand(N, device=DEV) * 0.01
        bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        med, best = timeit(lambda: ops.scaled_mm(a8, b8, sx, sw, bias, torch.bfloat16))
        line = {"ours_int8_ms": med, "ours_int8_tflops": fl / med / 1e9, "ours_int8_best_tflops": fl / best / 1e9}
        try:
            if M > 16:
                med_c, best_c = timeit(lambda: torch._int_mm(a8, b8.t()))
                line["cublaslt_int8_mm_ms"] = med_c
                line["cublaslt_int8_tflops"] = fl / med_c / 1e9
        except Exception as ex:  # noqa: BLE001
            line["cublaslt_int8_err"] = str(ex)[:80]
        af = a8.to(torch.float32).clamp(-8, 8).to(torch.float8_e4m3fn)
        bf = b8.to(torch.float32).clamp(-8, 8).to(torch.float8_e4m3fn)
        med, best = timeit(lambda: ops.scaled_mm(af, bf, sx, sw, bias, torch.bfloat16))
        line.update({"ours_fp8_ms": med, "ours_fp8_tflops": fl / med / 1e9})
        try:
            one = torch.ones(1, device=DEV)
            med_c, _ = timeit(lambda: torch._scaled_mm(af, bf.t(), scale_a=one, scale_b=one, out_dtype=torch.bfloat16))
            line["cublaslt_fp8_tflops"] = fl / med_c / 1e9
        except Exception as ex:  # noqa: BLE001
            line["cublaslt_fp8_err"] = str(ex)[:80]
        xb = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
        med, _ = timeit(lambda: ops.act_quant(xb, "int8"))
        line["act_quant_int8_ms"] = med
        line["act_quant_int8_gbs"] = 3.0 * M * K / med / 1e6
        med, _ = timeit(lambda: ops.act_quant(xb, "float8_e4m3fn", hadamard_group=256 if K % 256 == 0 else 128))
        line["act_quant_fp8_had_ms"] = med
        line["act_quant_fp8_had_gbs"] = 3.0 * M * K / med / 1e6
        res[f"{M}x{N}x{K}"] = line
        print(f"{M}x{N}x{K}", json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in line.items()}), flush=True)
        del a8, b8, af, bf, xb
    # dequant kernel (K3) and requant (K4) on SD-XL / FLUX weight shapes
    for (N, K) in [(10240, 1280), (1280, 5120), (12288, 3072), (4096, 4096)]:
        for wd, bits, gs in [("int4", 4, 128), ("int8", 8, K), ("uint3", 3, 64), ("float6_e3m2fn", 6, K)]:
            nbytes = N * K * bits // 8
            w = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device=DEV)
            if wd == "int8":
                w = w.view(torch.int8).view(N, K)
            scale = torch.rand(N * (K // gs), device=DEV) * 0.01
            zp = torch.rand(N * (K // gs), device=DEV) if wd.startswith("uint") else None
            med, _ = timeit(lambda: ops.dequant(w, wd, scale, zp, N, K, gs, torch.bfloat16))
            by = nbytes + 4 * N * (K // gs) * (2 if zp is not None else 1) + 2 * N * K
            print(f"dequant {wd} {N}x{K} g{gs}: {med:.4f} ms  {by / med / 1e6:.0f} GB/s", flush=True)
            res[f"dequant_{wd}_{N}x{K}"] = {"ms": med, "gbs": by / med / 1e6}
    json.dump(res, open("gpurun_out/kernels.json", "w"), indent=1)


if __name__ == "__main__":
    main()
