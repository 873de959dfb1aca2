"""Quantized conv forward (W8A8 over the im2col view) at SD-XL UNet conv shapes: K2c (gather quantiser), K1, the output permute,
the whole SDNQConv2d.forward, and the bf16 library convolution for scale.  CUDA-graph replay timed with CUDA events.
    python tools/conv_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import SDNQConfig, ops, sdnq_quantize_layer
from sdnq_b200.forward import matmul_operand

DEV = "cuda"


def graph_time(fn, reps=10):
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
    return t0.elapsed_time(t1) / reps * 1000


for (B, C, H, W, N, k, s, p) in [(1, 320, 128, 128, 320, 3, 1, 1), (1, 640, 64, 64, 640, 3, 1, 1), (1, 1280, 32, 32, 1280, 3, 1, 1),
                                 (1, 320, 128, 128, 320, 3, 2, 1), (1, 1920, 32, 32, 1280, 3, 1, 1), (1, 960, 64, 64, 640, 1, 1, 0), (2, 640, 64, 64, 640, 3, 1, 1)]:
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(C, N, k, stride=s, padding=p).to(DEV, torch.bfloat16)
    x = torch.randn(B, C, H, W, device=DEV, dtype=torch.bfloat16)
    layer, _ = sdnq_quantize_layer(torch.nn.Conv2d(C, N, k, stride=s, padding=p).to(DEV, torch.bfloat16),
                                   SDNQConfig(weights_dtype="int8", quant_conv=True, use_quantized_matmul=True, use_quantized_matmul_conv=True))
    op = matmul_operand(layer)
    xq, sx, _, _, _, (b_, ho, wo) = ops.conv_act_quant(x, (k, k), (s, s), (p, p), (1, 1), "int8")
    out = ops.scaled_mm(xq, op.wq, sx, op.sw, layer.bias, torch.bfloat16)
    M, K = xq.shape
    t_k2 = graph_time(lambda: ops.conv_act_quant(x, (k, k), (s, s), (p, p), (1, 1), "int8"))
    t_k1 = graph_time(lambda: ops.scaled_mm(xq, op.wq, sx, op.sw, layer.bias, torch.bfloat16))
    t_perm = graph_time(lambda: ops.rows_to_nchw(out, b_, ho * wo))
    t_all = graph_time(lambda: layer(x))
    t_lib = graph_time(lambda: conv(x))
    t_unf = graph_time(lambda: ops.act_quant(torch.nn.functional.unfold(x, kernel_size=k, padding=p, stride=s).transpose(1, 2).reshape(M, K).contiguous(), "int8"))
    fl = 2.0 * M * N * K
    print(f"conv {C}->{N} {k}x{k}/s{s} @{H}x{W}  M={M} K={K}:  K2c {t_k2:7.1f} us ({(M * K + 2 * B * C * H * W) / t_k2 / 1e3:6.0f} GB/s)  K1 {t_k1:7.1f} us ({fl / t_k1 / 1e6:6.0f} TF)  "
          f"permute {t_perm:6.1f}  forward {t_all:7.1f} us ({fl / t_all / 1e6:6.0f} TF)  |  unfold+K2 {t_unf:7.1f}  bf16 cuDNN conv {t_lib:7.1f} us", flush=True)
