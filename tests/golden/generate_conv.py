"""Generate the committed convolution fixtures (tests/golden/conv_*.npz) from the upstream reference.

Run in the authoring container only (needs /root/reference):

    SDNQ_USE_CONTIGUOUS_MM=0 SDNQ_ALLOW_FP8_MM=1 SDNQ_USE_TORCH_COMPILE=0 python tests/golden/generate_conv.py

Same conventions as generate.py (B200 flag state, bf16 as uint16 bit patterns, fp8 as uint8, K-major tensors stored as their
physical transpose under `key__T`).  Everything written is data produced by running the unmodified reference.
"""
import copy
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_loader import load_reference  # noqa: E402

assert os.environ.get("SDNQ_USE_CONTIGUOUS_MM") == "0" and os.environ.get("SDNQ_ALLOW_FP8_MM") == "1", __doc__
sdnq = load_reference()
from sdnq import SDNQConfig  # noqa: E402
from sdnq.quant_utils import get_hadamard, rotate_hadamard  # noqa: E402
from sdnq.quantizer import sdnq_quantize_layer  # noqa: E402
from sdnq.layers.conv.forward import get_conv_args, process_conv_input  # noqa: E402
from sdnq.layers.linear.linear_int8 import quantize_int_mm_input  # noqa: E402
from sdnq.layers.linear.linear_uint8 import quantize_uint_mm_input  # noqa: E402
from sdnq.layers.linear.linear_fp8 import quantize_fp_mm_input  # noqa: E402
from generate import tinfo, to_np  # noqa: E402

W8 = dict(quant_conv=True, use_quantized_matmul=True, use_quantized_matmul_conv=True)
DQ = dict(quant_conv=True)
CASES = {
    # name: (module ctor, ctor kwargs, input shape, SDNQConfig kwargs)
    "int8_3x3_w8a8":          ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1), (2, 32, 12, 10), dict(weights_dtype="int8", **W8)),
    "int8_1x1_w8a8_nobias":   ("Conv2d", dict(in_channels=64, out_channels=32, kernel_size=1, bias=False), (1, 64, 9, 8), dict(weights_dtype="int8", **W8)),
    "fp8_3x3_s2_w8a8":        ("Conv2d", dict(in_channels=32, out_channels=48, kernel_size=3, stride=2, padding=1), (2, 32, 13, 11), dict(weights_dtype="float8_e4m3fn", **W8)),
    "uint8_3x3_w8a8":         ("Conv2d", dict(in_channels=32, out_channels=32, kernel_size=3, padding=1), (1, 32, 8, 8), dict(weights_dtype="uint8", **W8)),
    "int8_5x3_dil_w8a8":      ("Conv2d", dict(in_channels=48, out_channels=32, kernel_size=(5, 3), stride=(1, 2), padding=(4, 1), dilation=(2, 1)), (1, 48, 10, 12), dict(weights_dtype="int8", **W8)),
    "int8_3x3_hadamard_w8a8": ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1), (1, 32, 8, 8), dict(weights_dtype="int8", use_hadamard=True, **W8)),
    "int8_3x3_svd_w8a8":      ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1), (1, 32, 8, 8), dict(weights_dtype="int8", use_svd=True, svd_rank=8, **W8)),
    "int8_conv1d_w8a8":       ("Conv1d", dict(in_channels=32, out_channels=48, kernel_size=3, padding=1), (3, 32, 40), dict(weights_dtype="int8", **W8)),
    "int8_3x3_w8a8_small_m":  ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1), (1, 32, 4, 4), dict(weights_dtype="int8", **W8)),
    "int8_3x3_dequant":       ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1), (2, 32, 8, 8), dict(weights_dtype="int8", **DQ)),
    "uint4_3x3_g16_dequant":  ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1), (2, 32, 8, 8), dict(weights_dtype="uint4", group_size=16, **DQ)),
    "int5_3x3_dequant":       ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1, bias=False), (2, 32, 8, 8), dict(weights_dtype="int5", **DQ)),
    "int8_3x3_svd_dequant":   ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1), (2, 32, 8, 8), dict(weights_dtype="int8", use_svd=True, svd_rank=8, **DQ)),
    "fp8_3x3_tensorwise_dequant": ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1), (2, 32, 8, 8), dict(weights_dtype="float8_e4m3fn", group_size=-2, **DQ)),
    "int8_convT_dequant":     ("ConvTranspose2d", dict(in_channels=32, out_channels=48, kernel_size=4, stride=2, padding=1), (2, 32, 6, 6), dict(weights_dtype="int8", **DQ)),
    "uint4_convT_g16_dequant": ("ConvTranspose2d", dict(in_channels=32, out_channels=32, kernel_size=3, stride=1, padding=1), (2, 32, 6, 6), dict(weights_dtype="uint4", group_size=16, **DQ)),
    "int8_conv1dT_dequant":   ("ConvTranspose1d", dict(in_channels=32, out_channels=64, kernel_size=4, stride=2, padding=1), (2, 32, 20), dict(weights_dtype="int8", **DQ)),
    "int8_conv3d_dequant":    ("Conv3d", dict(in_channels=32, out_channels=32, kernel_size=3, padding=1), (1, 32, 4, 6, 6), dict(weights_dtype="int8", **DQ)),
    # round 2: re-quantised (grouped / packed) conv weights on the quantized matmul, SVD factors of a re-quantised layer, padding modes,
    # Hadamard-rotated conv weights on the dequant path
    "uint4_3x3_g16_w8a8":     ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1), (1, 32, 8, 8), dict(weights_dtype="uint4", group_size=16, **W8)),
    "uint4_3x3_g16_svd_w8a8": ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1), (1, 32, 8, 8), dict(weights_dtype="uint4", group_size=16, use_svd=True, svd_rank=8, **W8)),
    "uint4_3x3_g16_svd_w8a8_small_m": ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1), (1, 32, 4, 4), dict(weights_dtype="uint4", group_size=16, use_svd=True, svd_rank=8, **W8)),
    "int5_3x3_g16_w8a8":      ("Conv2d", dict(in_channels=32, out_channels=48, kernel_size=3, padding=1), (2, 32, 8, 6), dict(weights_dtype="int5", group_size=16, **W8)),
    "float6_3x3_g16_fp8mm":   ("Conv2d", dict(in_channels=32, out_channels=48, kernel_size=3, padding=1), (2, 32, 8, 6), dict(weights_dtype="float6_e3m2fn", group_size=16, **W8)),
    "int8_3x3_reflect_w8a8":  ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1, padding_mode="reflect"), (2, 32, 9, 8), dict(weights_dtype="int8", **W8)),
    "int8_3x3_hadamard_dequant": ("Conv2d", dict(in_channels=32, out_channels=64, kernel_size=3, padding=1), (2, 32, 8, 8), dict(weights_dtype="int8", use_hadamard=True, **DQ)),
}


def run(name, cls, kw, xshape, cfg):
    torch.manual_seed(sum(map(ord, name)))
    mod = getattr(torch.nn, cls)(**kw).to(torch.bfloat16)
    x = torch.randn(*xshape).to(torch.bfloat16)
    arrays = {"x": to_np(x), "w_orig": to_np(mod.weight.detach().clone())}
    if mod.bias is not None:
        arrays["bias"] = to_np(mod.bias.detach().clone())
    layer = sdnq_quantize_layer(copy.deepcopy(mod), SDNQConfig(**cfg))[0]
    d = layer.sdnq_dequantizer
    meta = {"name": name, "module": cls, "module_kwargs": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()},
            "x_shape": list(xshape), "config": cfg, "tensors": {}}
    for key in ("weight", "scale", "zero_point", "svd_up", "svd_down"):
        t = getattr(layer, key)
        meta["tensors"][key] = tinfo(t)
        if t is not None:
            tt = t.detach()
            if tt.ndim == 2 and not tt.is_contiguous() and tt.t().is_contiguous():
                arrays[key + "__T"] = to_np(tt.t())
            else:
                arrays[key] = to_np(tt)
    meta["dequantizer"] = {k: (str(v).replace("torch.", "") if isinstance(v, torch.dtype) else (list(v) if isinstance(v, (torch.Size, tuple)) else v))
                           for k, v in d.__dict__.items()}
    meta["forward_func"] = layer.forward_func.__name__
    with torch.no_grad():
        y = layer(x)
        arrays["y"] = to_np(y)
        try:
            W = d(layer.weight, layer.scale, layer.zero_point, layer.svd_up, layer.svd_down, skip_quantized_matmul=d.use_quantized_matmul)
            arrays["w_dequant"] = to_np(W)
            meta["w_dequant"] = tinfo(W)
        except RuntimeError as ex:      # the reference itself cannot dequantise some combinations (e.g. Hadamard + matmul-layout conv)
            meta["w_dequant_error"] = str(ex)[:200]
        fn = layer.forward_func.__name__
        if fn.endswith("_matmul") and x.numel() / x.shape[2] >= 32:
            conv_type, stride, padding, dilation = get_conv_args(x.ndim, layer.stride, layer.padding, layer.dilation)
            cols, mm_shape = process_conv_input(conv_type, x, layer._reversed_padding_repeated_twice, layer.padding_mode, d.result_shape,
                                                stride, padding, dilation)
            meta["mm_output_shape"] = list(mm_shape)
            if d.use_hadamard:
                cols = rotate_hadamard(cols, hadamard=get_hadamard(d.hadamard_group_size, dtype=x.dtype, device=x.device))
                arrays["x_rot"] = to_np(cols.flatten(0, -2))
            arrays["cols"] = to_np(cols.flatten(0, -2)) if not d.use_hadamard else arrays["x_rot"]
            if fn.endswith("_int8_matmul"):
                xq, sx = quantize_int_mm_input(cols, dtype=layer.scale.dtype)
            elif fn.endswith("_uint8_matmul"):
                xq, sx, zx = quantize_uint_mm_input(cols, dtype=layer.scale.dtype)
                arrays["mm_zx"] = to_np(zx)
            else:
                xq, sx = quantize_fp_mm_input(cols, dtype=layer.scale.dtype)
            arrays["mm_xq"] = to_np(xq)
            arrays["mm_sx"] = to_np(sx)
    arrays["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, f"conv_{name}.npz"), **arrays)
    print(f"{name:30s} fwd={fn:42s} w={tuple(layer.weight.shape)} scale={tuple(layer.scale.shape)} qshape={tuple(d.quantized_weight_shape)} "
          f"result_shape={None if d.result_shape is None else tuple(d.result_shape)} gs={d.group_size} requant={d.re_quantize_for_matmul} y={tuple(y.shape)}")


if __name__ == "__main__":
    only = [n for n in os.environ.get("SDNQ_GOLDEN_ONLY", "").split(",") if n]      # regenerate just the named cases
    for name, (cls, kw, xshape, cfg) in CASES.items():
        if only and name not in only:
            continue
        try:
            run(name, cls, kw, xshape, cfg)
        except Exception as ex:      # combinations the reference itself cannot run produce no fixture
            print(f"{name:30s} REFERENCE FAILS: {type(ex).__name__}: {str(ex)[:120]}")
    print("bytes:", sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.startswith("conv_")))
