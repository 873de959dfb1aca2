"""W8A8 Linear with SVD rank-32 correction: K7 (svd_low) + rank-r accumulate inside K1 versus the reference-shaped two library GEMMs +
dense [M,N] bias (SDNQ_B200_SVD_FUSED=0).  `count` layers with their own weights (> L2) in one CUDA graph, replay timed with CUDA
events.   python tools/svd_w8a8_ab.py"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
from tools.shape_breakdown import graph_time

DEV = "cuda"
for wd, hadamard in (("int8", False), ("float8_e4m3fn", True), ("int4", False)):
    for (M, N, K) in [(1024, 1280, 1280), (4096, 640, 640), (4096, 5120, 640), (1024, 1280, 5120), (16384, 3072, 3072), (16384, 12288, 3072)]:
        count = max(2, min(24, int(300e6 // (N * K))))
        torch.manual_seed(0)
        base = torch.nn.Linear(K, N).to(torch.bfloat16)
        cfg = SDNQConfig(weights_dtype=wd, use_quantized_matmul=True, use_svd=True, svd_rank=32, use_hadamard=hadamard, group_size=-1 if wd == "int4" else 0)
        layer, _ = sdnq_quantize_layer(copy.deepcopy(base), cfg)
        layers = [copy.deepcopy(layer).to(DEV) for _ in range(count)]
        x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
        res, outs = {}, {}
        for tag, env in (("fused", "1"), ("dense-bias", "0")):
            os.environ["SDNQ_B200_SVD_FUSED"] = env
            keep = []

            def run():
                keep.clear()
                for l in layers:
                    keep.append(l(x))
            res[tag] = graph_time(run) / count * 1e3
            outs[tag] = keep[0].float()
        os.environ.pop("SDNQ_B200_SVD_FUSED", None)
        diff = (outs["fused"] - outs["dense-bias"]).abs().max().item() / outs["dense-bias"].abs().max().item()
        fl = 2 * M * N * K / 1e6
        print(f"{wd:>14s}{' +hadamard' if hadamard else '':10s} {M:6d}x{N:6d}x{K:5d} svd32: fused {res['fused']:8.2f} us ({fl / res['fused']:7.1f} TF)   "
              f"two GEMMs + [M,N] bias {res['dense-bias']:8.2f} us ({fl / res['dense-bias']:7.1f} TF)   max rel diff {diff:.2e}", flush=True)
