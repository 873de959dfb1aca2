"""Shared helpers for the parity tests: fixture -> torch tensors, bf16 ulp distance, oracle bridges."""
import copy
import glob
import json
import os

import numpy as np
import torch

from oracle import sdnq_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LAYER_FILES = sorted(glob.glob(os.path.join(GOLDEN, "layer_*.npz")))
LAYER_IDS = [os.path.basename(p)[6:-4] for p in LAYER_FILES]
CONV_FILES = sorted(glob.glob(os.path.join(GOLDEN, "conv_*.npz")))
CONV_IDS = [os.path.basename(p)[5:-4] for p in CONV_FILES]

_TORCH = {"bfloat16": torch.bfloat16, "float32": torch.float32, "float16": torch.float16, "int8": torch.int8, "uint8": torch.uint8,
          "int64": torch.int64, "int16": torch.int16, "float8_e4m3fn": torch.float8_e4m3fn, "float8_e5m2": torch.float8_e5m2, "bool": torch.bool,
          "int32": torch.int32}


def np_to_torch(a: np.ndarray, dtype_name: str, device="cpu") -> torch.Tensor:
    """inverse of tests/golden/generate.py:to_np (bf16 / fp8 are stored as bit patterns)."""
    if dtype_name == "bfloat16":
        t = torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)
    elif dtype_name in ("float8_e4m3fn", "float8_e5m2"):
        t = torch.from_numpy(a.copy()).view(_TORCH[dtype_name])
    else:
        t = torch.from_numpy(a.copy())
        assert t.dtype == _TORCH[dtype_name], (t.dtype, dtype_name)
    return t.to(device)


def fixture_tensors(path, device="cpu"):
    """-> (tensors dict with the reference's logical shapes/strides, raw arrays, meta)."""
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    out = {}
    for key in ("weight", "scale", "zero_point", "svd_up", "svd_down"):
        info = meta["tensors"][key]
        if info is None:
            out[key] = None
        elif key + "__T" in z.files:
            out[key] = np_to_torch(z[key + "__T"], info["dtype"], device).t()
        else:
            out[key] = np_to_torch(z[key], info["dtype"], device)
        if out[key] is not None:
            strides_ok = all(a == b or n == 1 for a, b, n in zip(out[key].stride(), info["stride"], info["shape"]))
            assert list(out[key].shape) == info["shape"] and strides_ok, (key, out[key].shape, out[key].stride(), info)
    out["x"] = np_to_torch(z["x"], "bfloat16", device)
    out["bias"] = np_to_torch(z["bias"], "bfloat16", device) if "bias" in z.files else None
    out["w_orig"] = np_to_torch(z["w_orig"], "bfloat16", device)
    return out, z, meta


def bf16_ulp_diff(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """distance in bf16 ulps between two bf16 tensors (monotone integer mapping of the bit patterns)."""
    def key(t):
        i = t.contiguous().view(torch.int16).to(torch.int32) & 0xFFFF
        return torch.where((i & 0x8000) != 0, -(i & 0x7FFF), i)
    return (key(a) - key(b)).abs()


def to_f32_np(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()


def oracle_layer_from_torch(tensors, dequantizer_meta) -> O.Layer:
    def conv(t):
        if t is None:
            return None
        if t.dtype in (torch.bfloat16, torch.float16, torch.float32, torch.float8_e4m3fn, torch.float8_e5m2):
            return t.detach().float().cpu().numpy()
        return t.detach().cpu().numpy()
    return O.Layer(conv(tensors["weight"]), conv(tensors["scale"]), conv(tensors["zero_point"]), conv(tensors["svd_up"]),
                   conv(tensors["svd_down"]), bias=conv(tensors.get("bias")), **dequantizer_meta)


def build_conv_layer(path):
    """quantise w_orig with our host code (checks it stores what the reference stores), keep the reference's tensors."""
    from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
    t, z, meta = fixture_tensors(path)
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in meta["module_kwargs"].items()}
    mod = getattr(torch.nn, meta["module"])(**kw).to(torch.bfloat16)
    with torch.no_grad():
        mod.weight.copy_(t["w_orig"])
        if t["bias"] is not None:
            mod.bias.copy_(t["bias"])
    layer, _ = sdnq_quantize_layer(copy.deepcopy(mod), SDNQConfig(**meta["config"]))
    ours = layer.sdnq_dequantizer
    ref = meta["dequantizer"]
    for key in ("group_size", "use_quantized_matmul", "re_quantize_for_matmul", "weights_dtype", "quantized_matmul_dtype", "use_hadamard"):
        assert getattr(ours, key) == ref[key], (key, getattr(ours, key), ref[key])
    assert list(ours.quantized_weight_shape) == ref["quantized_weight_shape"]
    assert (None if ours.result_shape is None else list(ours.result_shape)) == ref["result_shape"]
    assert layer.forward_func.__name__ == meta["forward_func"]
    for key in ("weight", "scale", "zero_point", "svd_up", "svd_down"):
        mine = getattr(layer, key)
        assert (mine is None) == (t[key] is None), key
        if t[key] is not None:
            assert mine.shape == t[key].shape, (key, mine.shape, t[key].shape)
            if key in ("weight", "scale", "zero_point") and t["svd_up"] is None:
                a, b = (v.view(torch.uint8) if v.dtype in (torch.float8_e4m3fn, torch.float8_e5m2) else v for v in (mine, t[key]))
                assert torch.equal(a, b), f"stored {key} differs from the reference's"
            setattr(layer, key, torch.nn.Parameter(t[key], requires_grad=False))
    return layer, t, z, meta


