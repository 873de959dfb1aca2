"""Model-level parity fixture: run the reference's sdnq_post_load_quant on tests/golden/toy_model.py for a few configs and record,
for every module / state-dict entry, what came out (class names, forward function names, dequantizer metadata, tensor shapes /
dtypes / strides and a hash of the bytes) plus the resulting quantization_config lists.

    SDNQ_USE_CONTIGUOUS_MM=0 SDNQ_ALLOW_FP8_MM=1 SDNQ_USE_TORCH_COMPILE=0 python tests/golden/generate_model.py
"""
import hashlib
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_loader import load_reference  # noqa: E402

assert os.environ.get("SDNQ_USE_CONTIGUOUS_MM") == "0" and os.environ.get("SDNQ_ALLOW_FP8_MM") == "1", __doc__
load_reference()
from sdnq import sdnq_post_load_quant  # noqa: E402
import toy_model  # noqa: E402


def describe(model):
    mods = {}
    for name, m in model.named_modules():
        if name == "":
            continue
        e = {"class": type(m).__name__}
        d = getattr(m, "sdnq_dequantizer", None)
        if d is not None:
            e["forward_func"] = m.forward_func.__name__
            e["dequantizer"] = {k: (str(v).replace("torch.", "") if isinstance(v, torch.dtype) else (list(v) if isinstance(v, (torch.Size, tuple)) else v))
                                for k, v in d.__dict__.items()}
        mods[name] = e
    tensors = {}
    for key, t in model.state_dict().items():
        tt = t.detach()
        phys = tt.t() if (tt.ndim == 2 and not tt.is_contiguous() and tt.t().is_contiguous()) else tt.contiguous()
        raw = phys.view(torch.uint8) if phys.dtype in (torch.float8_e4m3fn, torch.float8_e5m2) else phys
        data = raw.view(torch.uint8).numpy().tobytes() if raw.dtype != torch.bfloat16 else raw.view(torch.int16).numpy().tobytes()
        tensors[key] = {"dtype": str(tt.dtype).replace("torch.", ""), "shape": list(tt.shape), "stride": list(tt.stride()),
                        "sha1": hashlib.sha1(data).hexdigest()}
    return mods, tensors


out = {}
for name, cfg in toy_model.CONFIGS.items():
    model = toy_model.build()
    model = sdnq_post_load_quant(model, **cfg)
    mods, tensors = describe(model)
    qc = model.quantization_config
    out[name] = {"config": cfg, "modules": mods, "tensors": tensors,
                 "modules_to_not_convert": sorted(qc.modules_to_not_convert), "modules_dtype_dict": {k: sorted(v) for k, v in qc.modules_dtype_dict.items()},
                 "modules_to_not_use_matmul": sorted(qc.modules_to_not_use_matmul)}
    nq = sum(1 for e in mods.values() if "forward_func" in e)
    print(f"{name:24s} quantised modules: {nq:2d}  state-dict entries: {len(tensors)}")
json.dump(out, open(os.path.join(HERE, "model_parity.json"), "w"), indent=0, sort_keys=True)
print("bytes:", os.path.getsize(os.path.join(HERE, "model_parity.json")))

# ---- apply_sdnq_options_to_model: in-place layout / option flips on an already quantised model (loader.py:221-346)
from sdnq.loader import apply_sdnq_options_to_model  # noqa: E402

FLIPS = {
    "int8_off_to_on": (dict(weights_dtype="int8"), dict(use_quantized_matmul=True)),
    "int8_on_to_off": (dict(weights_dtype="int8", use_quantized_matmul=True), dict(use_quantized_matmul=False)),
    "fp8_hadamard_off_to_on": (dict(weights_dtype="float8_e4m3fn", use_hadamard=True), dict(use_quantized_matmul=True)),
    "uint4_off_to_on": (dict(weights_dtype="uint4"), dict(use_quantized_matmul=True)),
    "int8_svd_on_to_off": (dict(weights_dtype="int8", use_quantized_matmul=True, use_svd=True, svd_rank=8), dict(use_quantized_matmul=False)),
    "int8_conv_off_to_on": (dict(weights_dtype="int8", quant_conv=True), dict(use_quantized_matmul=True)),
}
flips = {}
for name, (cfg, opts) in FLIPS.items():
    torch.manual_seed(1)
    model = sdnq_post_load_quant(toy_model.build(), **cfg)
    before_mods, before = describe(model)
    model = apply_sdnq_options_to_model(model, **opts)
    mods, tensors = describe(model)
    flips[name] = {"config": cfg, "options": opts, "modules": mods, "tensors": tensors, "tensors_before": before}
    print(f"flip {name:24s} changed entries: {sum(1 for k in tensors if tensors[k] != before.get(k))}")
json.dump(flips, open(os.path.join(HERE, "model_flips.json"), "w"), indent=0, sort_keys=True)

# ---- dynamic quantisation (quantizer.py:280-419): the per-layer dtype search of sdnq_quantize_layer_weight_dynamic
DYNAMIC = {
    "uint4_dynamic": dict(weights_dtype="uint4", use_dynamic_quantization=True),
    "int3_dynamic_tight": dict(weights_dtype="int3", use_dynamic_quantization=True, dynamic_loss_threshold=3e-3),
    "uint4_dynamic_w8a8": dict(weights_dtype="uint4", use_dynamic_quantization=True, use_quantized_matmul=True, dynamic_loss_threshold=2e-3),
}
dyn = {}
for name, cfg in DYNAMIC.items():
    model = sdnq_post_load_quant(toy_model.build(), **cfg)
    mods, tensors = describe(model)
    qc = model.quantization_config
    dyn[name] = {"config": cfg, "modules": mods, "tensors": tensors, "modules_to_not_convert": sorted(qc.modules_to_not_convert),
                 "modules_dtype_dict": {k: sorted(v) for k, v in qc.modules_dtype_dict.items()},
                 "modules_to_not_use_matmul": sorted(qc.modules_to_not_use_matmul)}
    print(f"dynamic {name:22s} ->", {k: len(v) for k, v in qc.modules_dtype_dict.items()})
json.dump(dyn, open(os.path.join(HERE, "model_dynamic.json"), "w"), indent=0, sort_keys=True)

# ---- get_forward_func dispatch table (forward.py:6-57)
from sdnq.forward import get_forward_func  # noqa: E402

table = {}
for c in ["Linear", "SDNQLinear", "Conv1d", "Conv2d", "Conv3d", "SDNQConv2d", "ConvTranspose1d", "ConvTranspose2d", "ConvTranspose3d", "Embedding"]:
    for m in ["int8", "uint8", "float8_e4m3fn", "float16"]:
        for use in (False, True):
            table[f"{c}|{m}|{int(use)}"] = get_forward_func(c, m, use).__name__
json.dump(table, open(os.path.join(HERE, "forward_dispatch.json"), "w"), indent=0, sort_keys=True)

# ---- SDNQConfig.to_dict() (what quantization_config.json carries, quantizer.py:1075-1079)
from sdnq import SDNQConfig  # noqa: E402

cases = dict(toy_model.CONFIGS)
cases.update({
    "defaults": {},
    "int4_svd": dict(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32),
    "dynamic": dict(weights_dtype="uint4", use_dynamic_quantization=True, dynamic_loss_threshold=0.01),
    "fp8_tensorwise": dict(weights_dtype="float8_e4m3fn", group_size=-2, dequantize_fp32=False),
    "matmul_dtype": dict(weights_dtype="int6", quantized_matmul_dtype="float8_e4m3fn", use_quantized_matmul=True),
    "codebook": dict(weights_dtype="uint4", use_codebook=True, codebook_steps=12),
})
json.dump({n: {"kwargs": c, "to_dict": json.loads(json.dumps(SDNQConfig(**c).to_dict(), default=str))} for n, c in cases.items()},
          open(os.path.join(HERE, "config_dicts.json"), "w"), indent=0, sort_keys=True)

# ---- checkpoints of reference-quantised toy models: what save_pretrained writes (every tensor .contiguous() into safetensors) plus
#      the config's to_dict() as quantization_config.json (QuantizationConfigMixin.to_json_file)
from safetensors.torch import save_file  # noqa: E402

for name in ("uint4_conv_embedding", "int8_w8a8_conv"):
    model = sdnq_post_load_quant(toy_model.build(), **toy_model.CONFIGS[name])
    ckpt = os.path.join(HERE, "ckpt_" + name)
    os.makedirs(ckpt, exist_ok=True)
    save_file({k: v.contiguous() for k, v in model.state_dict().items()}, os.path.join(ckpt, "model.safetensors"))
    with open(os.path.join(ckpt, "quantization_config.json"), "w") as f:
        json.dump(json.loads(json.dumps(model.quantization_config.to_dict(), default=str)), f, indent=2, sort_keys=True)
print("dispatch / config / checkpoint fixtures written")
