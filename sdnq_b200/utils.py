"""Host-side policy helpers of the SDNQ surface: module-name matching, eligibility rules, per-layer kwargs.

Behaviour follows reference utils.py:14-249 (the rules decide which tensors exist in a checkpoint, so they are part
of the drop-in boundary); the implementation is ours."""
import fnmatch
import re

import torch

from .common import (allowed_types, common_skip_keys, conv_transpose_types, conv_types, dtype_dict, embedding_types,
                     module_skip_keys_dict)


def is_pow2(n: int) -> bool:
    return n & (n - 1) == 0


def is_pow4(n: int) -> bool:
    # a power of two whose single set bit sits at an even position
    return is_pow2(n) and n.bit_length() % 2 == 1


def next_power_of_2(n: int) -> int:
    return n if is_pow2(n) else 1 << n.bit_length()


def check_param_name_in(param_name: str, param_list) -> str | None:
    """First key of `param_list` that selects `param_name`, else None (reference utils.py:56-70).

    A key starting with "." is a *prefix of the full dotted name*; any other key matches if it equals the name, equals one
    dotted segment, or -- when it contains "*" -- glob-matches from the start of the name (".*" meaning "any suffix")."""
    segments = param_name.split(".")
    for key in param_list:
        if key.startswith("."):
            if param_name.startswith(key[1:]):
                return key
            continue
        if key == param_name or key in segments:
            return key
        if "*" in key:
            pattern = key.replace(".*", "\\.*").replace("*", ".*")
            if re.match(pattern, param_name):
                return key
    return None


def check_quant_is_allowed(layer_class_name: str, weight: torch.Tensor, quantization_config, pre_quantized: bool = False) -> bool:
    """Is this layer quantised at all (reference utils.py:73-90)."""
    if layer_class_name not in allowed_types or weight.dtype not in (torch.float64, torch.float32, torch.float16, torch.bfloat16):
        return False
    if layer_class_name in embedding_types and not quantization_config.quant_embedding:
        return False
    is_conv = layer_class_name in conv_types or layer_class_name in conv_transpose_types
    if is_conv and not quantization_config.quant_conv:
        return False
    if pre_quantized:
        return True
    if layer_class_name in conv_types:
        channels = weight.shape[1]
    elif layer_class_name in conv_transpose_types:
        channels = weight.shape[0]
    else:
        channels = weight.shape[-1]
    return channels >= quantization_config.minimum_allowed_channel_size and weight.numel() >= quantization_config.minimum_allowed_numel


def check_quantized_matmul_is_allowed(use_quantized_matmul: bool, output_channel_size: int, channel_size: int) -> bool:
    """W8A8 needs both dims >= 32 and multiples of 16 (reference utils.py:93-98)."""
    return bool(use_quantized_matmul and min(output_channel_size, channel_size) >= 32
                and output_channel_size % 16 == 0 and channel_size % 16 == 0)


_RUNTIME_ONLY_KEYS = ("is_integer", "quant_method", "quantization_device", "return_device", "non_blocking", "add_skip_keys",
                      "use_dynamic_quantization", "use_static_quantization", "use_stochastic_rounding", "use_grad_ckpt",
                      "is_training", "sdnq_version")


def get_quant_args_from_config(quantization_config) -> dict:
    """Config -> the kwargs that describe the *format* (drops run-time-only switches; reference utils.py:101-122)."""
    cfg = quantization_config.to_dict() if hasattr(quantization_config, "to_dict") else dict(quantization_config)
    for key in _RUNTIME_ONLY_KEYS:
        cfg.pop(key, None)
    sub = cfg.get("modules_quant_config")
    if sub is not None:
        for key in sub:
            sub[key] = get_quant_args_from_config(sub[key])
    return cfg


def get_minimum_dtype(weights_dtype: str, param_name: str, modules_dtype_dict: dict) -> str:
    """Per-module dtype override; keys like "minimum_6bit" raise the width only if needed (reference utils.py:125-146)."""
    for key, names in modules_dtype_dict.items():
        if check_param_name_in(param_name, names) is None:
            continue
        key = key.lower()
        if not (key.startswith("minimum") or key.endswith(("bit", "bits"))):
            return key
        spec = key.removeprefix("minimum").removeprefix("-").removeprefix("_")
        spec = spec.removesuffix("bits").removesuffix("bit").removesuffix("-").removesuffix("_")
        unsigned = spec.startswith("uint")
        spec = spec.removeprefix("uint") if unsigned else spec.removeprefix("int")
        if dtype_dict[weights_dtype]["num_bits"] < int(spec):
            return ("uint" if unsigned or int(spec) <= 4 else "int") + spec
    return weights_dtype


def get_quantized_matmul_dtype(weights_dtype: str, quantized_matmul_dtype: str | None = None) -> str:
    """Default matmul dtype for a weight dtype (reference utils.py:203-214)."""
    if quantized_matmul_dtype is not None:
        return quantized_matmul_dtype
    info = dtype_dict[weights_dtype]
    if info["is_integer"]:
        return "uint8" if weights_dtype == "uint8" else "int8"
    return "float8_e4m3fn" if info["num_bits"] < 16 else "float16"


_PER_LAYER_FIELDS = ("weights_dtype", "quantized_matmul_dtype", "hadamard_group_size", "group_size", "svd_rank", "svd_steps",
                     "codebook_steps", "dynamic_loss_threshold", "use_svd", "use_hadamard", "use_codebook", "use_quantized_matmul",
                     "use_quantized_matmul_conv", "use_dynamic_quantization", "use_stochastic_rounding", "dequantize_fp32",
                     "non_blocking", "quantization_device", "return_device")


def get_quant_kwargs(layer: torch.nn.Module, quantization_config, torch_dtype: torch.dtype | None = None, param_name: str = "", **overrides) -> dict:
    """Resolve the quantisation kwargs of one layer: config -> explicit overrides -> modules_quant_config ->
    conv/linear matmul switch -> modules_dtype_dict -> modules_to_not_use_matmul (reference utils.py:149-200)."""
    from .config import SDNQConfig
    if not isinstance(quantization_config, SDNQConfig):
        quantization_config = SDNQConfig(**quantization_config)
    layer_class_name = layer.__class__.__name__
    kw = {name: getattr(quantization_config, name) for name in _PER_LAYER_FIELDS}
    kw.update(layer_class_name=layer_class_name, torch_dtype=torch_dtype, param_name=param_name)
    kw.update(overrides)
    key = check_param_name_in(kw["param_name"], quantization_config.modules_quant_config.keys())
    if key is not None:
        kw.update(quantization_config.modules_quant_config[key])
    conv_flag = kw.pop("use_quantized_matmul_conv")
    if layer_class_name in conv_types or layer_class_name in conv_transpose_types:
        kw["use_quantized_matmul"] = conv_flag
    if not kw["use_dynamic_quantization"]:
        kw.pop("dynamic_loss_threshold")
    kw["weights_dtype"] = get_minimum_dtype(kw["weights_dtype"], kw["param_name"], quantization_config.modules_dtype_dict)
    if check_param_name_in(kw["param_name"], quantization_config.modules_to_not_use_matmul) is not None:
        kw["use_quantized_matmul"] = False
    return kw


def add_module_skip_keys(model: torch.nn.Module, quantization_config):
    """Extend the config with the model's own fp32 / tied keys and the per-architecture policy table
    (reference utils.py:217-249)."""
    cfg = quantization_config
    keep_fp32 = getattr(model, "_keep_in_fp32_modules", None)
    if keep_fp32 is not None:
        cfg.modules_to_not_convert.extend(keep_fp32)
    tied = getattr(model, "_tied_weights_keys", None)
    if tied is not None:
        if isinstance(tied, dict):
            cfg.modules_to_not_convert.extend(tied.keys())
            cfg.modules_to_not_convert.extend(tied.values())
        else:
            cfg.modules_to_not_convert.extend(tied)
    policy = module_skip_keys_dict.get(model.__class__.__name__)
    if policy is None:
        cfg.modules_to_not_convert.extend(common_skip_keys)
        patterns = getattr(model, "_skip_layerwise_casting_patterns", None)
        if patterns is not None:
            cfg.modules_to_not_convert.extend(patterns)
    else:
        skip, dtype_overrides, no_matmul = policy
        cfg.modules_to_not_convert.extend(skip)
        for key, names in dtype_overrides.items():
            cfg.modules_dtype_dict.setdefault(key, [])
            cfg.modules_dtype_dict[key] = list(cfg.modules_dtype_dict[key]) + list(names)
        mm_dtype = get_quantized_matmul_dtype(cfg.weights_dtype, cfg.quantized_matmul_dtype)
        cfg.modules_to_not_use_matmul.extend(no_matmul.get(mm_dtype, []))
    cfg.modules_to_not_convert = list(set(cfg.modules_to_not_convert))
    cfg.modules_to_not_use_matmul = list(set(cfg.modules_to_not_use_matmul))
    for key in cfg.modules_dtype_dict:
        cfg.modules_dtype_dict[key] = list(set(cfg.modules_dtype_dict[key]))
    return model, cfg
