"""Per-shape cost of the W8A8 Linear inside a graph-replayed step (GPU box).

    python tools/shape_breakdown.py [sdxl|flux] [--fp8] [--hadamard G]

For every distinct (M, N, K) of the workload: `count` layers with their own weights (total footprint > L2, so weights
stream from HBM as in the real step) are captured as one CUDA graph -- K2+K1 pairs, K1 alone, K2 alone, and the
cuBLASLt int8 GEMM (torch._int_mm, no epilogue) for comparison -- and the replay is timed with CUDA events.
Prints microseconds per launch and the weighted total for the step.  Not part of the product."""
import json
import os
import sys
from collections import Counter

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import bench
from sdnq_b200 import ops

DEV = "cuda"


def graph_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        g.replay()
    t1.record()
    t1.synchronize()
    return t0.elapsed_time(t1) / reps


def main():
    which = "flux" if "flux" in sys.argv else "sdxl"
    fp8 = "--fp8" in sys.argv
    hg = int(sys.argv[sys.argv.index("--hadamard") + 1]) if "--hadamard" in sys.argv else 0
    layers = bench.flux_linears() if which == "flux" else bench.sdxl_linears()
    shapes = Counter((m, n, k) for _, m, n, k in layers if m >= 32)
    mm = "float8_e4m3fn" if fp8 else "int8"
    rows, tot = [], Counter()
    for (M, N, K), cnt in sorted(shapes.items()):
        count = max(4, min(64, int(400e6 // (N * K)) + 1))
        if which == "flux":
            count = min(count, 8)
        ws = []
        for _ in range(count):
            w = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=DEV)
            ws.append(w.to(torch.float32).clamp(-8, 8).to(torch.float8_e4m3fn) if fp8 else w)
        x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
        sw = torch.rand(N, device=DEV) * 0.01 + 0.001
        bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)
        xq, sx, *_ = ops.act_quant(x, mm, hadamard_group=hg)

        def pair():
            for w in ws:
                ops.linear_w8a8(x, w, mm, sw, bias=bias, hadamard_group=hg, out_dtype=torch.bfloat16)

        def fused():
            for w in ws:
                ops.linear_w8a8(x, w, mm, sw, bias=bias, out_dtype=torch.bfloat16, fused=True)

        def k1():
            for w in ws:
                ops.scaled_mm(xq, w, sx, sw, bias, torch.bfloat16)

        def k2():
            for _ in ws:
                ops.act_quant(x, mm, hadamard_group=hg)

        t_pair, t_k1, t_k2 = (1e3 * graph_time(f) / count for f in (pair, k1, k2))
        t_fused = float("nan")
        if not hg:
            try:
                t_fused = 1e3 * graph_time(fused) / count
            except Exception as ex:  # noqa: BLE001  (shape outside the fused kernel's coverage)
                print("   fused:", str(ex)[:100])
        t_lib = float("nan")
        if not fp8:
            def lib():
                for w in ws:
                    torch._int_mm(xq, w.t())
            try:
                t_lib = 1e3 * graph_time(lib) / count
            except Exception:  # noqa: BLE001
                pass
        fl = 2.0 * M * N * K
        rows.append(dict(M=M, N=N, K=K, layers=cnt, fused_us=t_fused, pair_us=t_pair, k1_us=t_k1, k2_us=t_k2, lib_us=t_lib, k1_tflops=fl / t_k1 / 1e6))
        tot["fused"] += cnt * t_fused
        tot["pair"] += cnt * t_pair
        tot["k1"] += cnt * t_k1
        tot["k2"] += cnt * t_k2
        tot["lib"] += cnt * t_lib
        print(f"M={M:6d} N={N:6d} K={K:6d} x{cnt:4d}: fused {t_fused:8.2f} us  pair {t_pair:8.2f}  K1 {t_k1:8.2f}  K2 {t_k2:7.2f}  cuBLASLt {t_lib:8.2f}   K1 {fl / t_k1 / 1e6:7.1f} TF",
              flush=True)
        del ws
        torch.cuda.empty_cache()
    print("step totals (ms): " + "  ".join(f"{k} {v / 1e3:.3f}" for k, v in tot.items()))
    os.makedirs("gpurun_out", exist_ok=True)
    tag = f"{which}_{'fp8' if fp8 else 'int8'}" + (f"_h{hg}" if hg else "") + os.environ.get("BREAKDOWN_TAG", "")
    json.dump(dict(rows=rows, totals_us=dict(tot)), open(f"gpurun_out/shape_breakdown_{tag}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
