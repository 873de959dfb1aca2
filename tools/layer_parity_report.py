"""Achieved error of every reference layer fixture through the public forward on the GPU: max and rms error relative to the output's
maximum, and the bf16-ulp statistics -- the numbers behind the tolerances of tests/test_layers_gpu.py.   python tools/layer_parity_report.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests.test_layers_gpu import build_layer
from tests.util import LAYER_FILES, LAYER_IDS, bf16_ulp_diff, np_to_torch

print("| fixture | path | max err / max|y| | rms err / max|y| | outputs off by 1 bf16 ulp | off by more |")
print("|---|---|---:|---:|---:|---:|")
for path, name in zip(LAYER_FILES, LAYER_IDS):
    layer, t, z, meta = build_layer(path)
    d = meta["dequantizer"]
    y = layer(t["x"].to("cuda"))
    yref = np_to_torch(z["y"], "bfloat16", "cuda")
    fin = torch.isfinite(yref)
    scale = float(yref[fin].float().abs().max())
    err = (y.float() - yref.float())[fin].abs()
    du = bf16_ulp_diff(y, yref)[fin]
    kind = ("W8A8" if d["use_quantized_matmul"] and meta["M"] >= 32 else "small-M" if meta["M"] < 32 else "dequant") + (" +rot" if d["use_hadamard"] else "") + (" +svd" if t["svd_up"] is not None else "")
    print(f"| {name} | {kind} | {float(err.max()) / scale:.2e} | {float(err.pow(2).mean().sqrt()) / scale:.2e} | {float((du == 1).float().mean()):.3%} | {float((du > 1).float().mean()):.3%} |")
