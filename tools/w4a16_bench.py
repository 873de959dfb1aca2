"""K6 (W4A16 dequant-path Linear) per-shape timing against the reference-shaped path of this library (K3 / K3s dequant + library
bf16 GEMM) and the bare library GEMM on an already dequantised weight.  Distinct layers rotate through more weight bytes than L2
holds; the launches are captured into one CUDA graph and the replay is timed with CUDA events.
    python tools/w4a16_bench.py [sdxl|flux] [nosvd]"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import SDNQConfig, sdnq_quantize_layer

DEV = "cuda"
SDXL = [(77, 1280, 2048), (1024, 1280, 1280), (1024, 1280, 5120), (1024, 10240, 1280), (4096, 640, 640), (4096, 640, 2560), (4096, 5120, 640)]
FLUX = [(16384, 3072, 3072), (16384, 12288, 3072), (16384, 3072, 12288)]


def graph_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
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
    which = sys.argv[1] if len(sys.argv) > 1 else "sdxl"
    svd = "nosvd" not in sys.argv
    cfg = dict(weights_dtype="int4", group_size=128, **(dict(use_svd=True, svd_rank=32, svd_steps=2) if svd else {}))
    for (M, N, K) in (SDXL if which == "sdxl" else FLUX):
        copies = max(2, min(12, int(200e6 // (N * K // 2)) + 1))
        torch.manual_seed(0)
        layers = []
        for _ in range(copies):
            lin = torch.nn.Linear(K, N, bias=True, device=DEV, dtype=torch.bfloat16)
            layers.append(sdnq_quantize_layer(lin, SDNQConfig(**cfg))[0])
        x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
        dense = [l.sdnq_dequantizer(l.weight, l.scale, l.zero_point, l.svd_up, l.svd_down) for l in layers[:4]]
        n = max(copies, 8)
        res = {}

        def run():
            for i in range(n):
                layers[i % copies](x)
        os.environ["SDNQ_B200_W4A16"] = "1"
        res["w4a16"] = graph_time(run) * 1000 / n
        os.environ["SDNQ_B200_W4A16"] = "0"
        res["dequant+gemm"] = graph_time(run) * 1000 / n
        os.environ["SDNQ_B200_DEQUANT_STREAM"] = "0"
        res["dequant+gemm(1 stream)"] = graph_time(run) * 1000 / n
        os.environ.pop("SDNQ_B200_DEQUANT_STREAM")
        os.environ.pop("SDNQ_B200_W4A16")

        def run_lib():
            for i in range(n):
                torch.nn.functional.linear(x, dense[i % len(dense)], layers[0].bias)
        res["bf16 gemm only"] = graph_time(run_lib) * 1000 / n
        fl = 2.0 * M * N * K
        print(f"M={M:6d} N={N:6d} K={K:6d} {'svd32' if svd else 'nosvd'}: " + "  ".join(f"{t} {us:8.2f} us ({fl / us / 1e6:6.1f} TF)" for t, us in res.items()), flush=True)


if __name__ == "__main__":
    main()
