"""K1 (scaled GEMM) A/B timing: single-CTA MMAs vs CTA pairs (tcgen05 cta_group::2), per shape, weights rotating through more
than L2, launches captured into one CUDA graph and the replay timed with CUDA events.
    python tools/gemm_ab.py [sdxl|flux|all] [fp8]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import ops

DEV = "cuda"
SDXL = [(77, 1280, 2048), (1024, 1280, 1280), (1024, 1280, 5120), (1024, 10240, 1280), (4096, 640, 640), (4096, 640, 2560), (4096, 5120, 640)]
FLUX = [(16384, 3072, 3072), (16384, 12288, 3072), (16384, 3072, 12288), (18432, 3072, 15360), (2048, 3072, 3072), (8192, 8192, 8192)]


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
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    fp8 = "fp8" in sys.argv
    shapes = SDXL if which == "sdxl" else FLUX if which == "flux" else SDXL + FLUX
    for (M, N, K) in shapes:
        copies = max(2, min(16, int(300e6 // (N * K)) + 1))
        a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=DEV)
        ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=DEV) for _ in range(copies)]
        if fp8:
            a = (a.float() / 16).to(torch.float8_e4m3fn)
            ws = [(w.float() / 16).to(torch.float8_e4m3fn) for w in ws]
        sx = torch.rand(M, device=DEV) * 0.01
        sw = torch.rand(N, device=DEV) * 0.01
        bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)
        layers = max(copies, 8)
        res = {}
        for tag, env in [("single", {"SDNQ_B200_CG": "1"}), ("pair128", {"SDNQ_B200_CG": "2", "SDNQ_B200_BN": "128"}),
                         ("pair256", {"SDNQ_B200_CG": "2", "SDNQ_B200_BN": "256"})]:
            for k in ("SDNQ_B200_CG", "SDNQ_B200_BN"):
                os.environ.pop(k, None)
            os.environ.update(env)

            def run():
                for i in range(layers):
                    ops.scaled_mm(a, ws[i % copies], sx, sw, bias, torch.bfloat16)
            ms = graph_time(run)
            res[tag] = ms * 1000 / layers
        for k in ("SDNQ_B200_CG", "SDNQ_B200_BN"):
            os.environ.pop(k, None)
        lib = None
        if not fp8 and M > 16:
            def run_lib():
                for i in range(layers):
                    torch._int_mm(a, ws[i % copies].t())
            lib = graph_time(run_lib) * 1000 / layers
        fl = 2.0 * M * N * K
        print(f"M={M:6d} N={N:6d} K={K:6d} {'fp8' if fp8 else 'int8'}: " + "  ".join(f"{t} {us:8.2f} us ({fl / us / 1e6:7.1f} TF)" for t, us in res.items())
              + (f"  cuBLASLt {lib:8.2f} us ({fl / lib / 1e6:7.1f} TF)" if lib else ""), flush=True)


if __name__ == "__main__":
    main()
