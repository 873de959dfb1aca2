"""Timing experiments on K6 (SDNQ_B200_W4A16_DBG bit mask: 1 no dequantise work, 2 no proxy fence, 4 no MMAs): which stage of the
per-k-block chain sets the pace.  Results are wrong by construction; only the times mean anything."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
from tools.w4a16_bench import graph_time

for svd in (False, True):
    cfg = dict(weights_dtype="int4", group_size=128, **(dict(use_svd=True, svd_rank=32, svd_steps=2) if svd else {}))
    for (M, N, K) in [(1024, 1280, 1280), (1024, 1280, 5120), (4096, 5120, 640)]:
        torch.manual_seed(0)
        layers = [sdnq_quantize_layer(torch.nn.Linear(K, N, bias=True, device="cuda", dtype=torch.bfloat16), SDNQConfig(**cfg))[0] for _ in range(8)]
        x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        out = []
        for dbg in (0, 1, 2, 3, 4, 5, 7):
            os.environ["SDNQ_B200_W4A16_DBG"] = str(dbg)

            def run():
                for l in layers:
                    l(x)
            out.append(f"dbg{dbg} {graph_time(run) * 1000 / len(layers):7.2f}")
        os.environ.pop("SDNQ_B200_W4A16_DBG")
        print(f"M={M} N={N} K={K} {'svd32' if svd else 'nosvd'} (us):  " + "  ".join(out), flush=True)
