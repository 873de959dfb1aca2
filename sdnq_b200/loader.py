"""Run-time option switches and (de)serialisation helpers of the SDNQ surface (reference loader.py).

`apply_sdnq_options_to_model` is the call users make to turn the W8A8 matmul on or off / change dtypes on an already
quantised model; it re-lays the stored tensors in place (transposes weight / scale / zero-point / SVD factors) and swaps
`forward_func`, so it is part of the drop-in boundary of the hot path.  Pure host / layout work."""
import json
import os

import torch

from .common import dtype_dict, linear_types
from .config import SDNQConfig
from .forward import get_forward_func
from .quant_math import prepare_svd_for_matmul, prepare_weight_for_matmul
from .quantizer import IS_FP8_MM_SUPPORTED, USE_TENSORWISE_FP8_MATMUL, sdnq_post_load_quant
from .utils import check_param_name_in


def _sdnq_children(model: torch.nn.Module, prefix: str = ""):
    for name, child in model.named_children():
        full = f"{prefix}.{name}" if prefix else name
        if hasattr(child, "sdnq_dequantizer"):
            yield model, name, child, full + ".weight"
        else:
            yield from _sdnq_children(child, full)


def post_process_model(model: torch.nn.Module) -> torch.nn.Module:
    """After tensors were assigned from a checkpoint: freeze them and restore the matmul layouts (reference loader.py:199-218)."""
    for _, _, layer, _ in _sdnq_children(model):
        deq = layer.sdnq_dequantizer
        layer.weight.requires_grad_(False)
        layer.scale.requires_grad_(False)
        if layer.zero_point is not None:
            layer.zero_point.requires_grad_(False)
        if deq.use_quantized_matmul and not deq.re_quantize_for_matmul:
            layer.weight.data = prepare_weight_for_matmul(layer.weight, matmul_dtype=deq.quantized_matmul_dtype)
        if layer.svd_up is not None:
            layer.svd_up.requires_grad_(False)
            layer.svd_down.requires_grad_(False)
            layer.svd_up.data, layer.svd_down.data = prepare_svd_for_matmul(layer.svd_up, layer.svd_down, deq.use_quantized_matmul)
    return model


def _wanted_matmul(layer, param_name, cfg, requested):
    """None = leave as is; True / False = target state (reference loader.py:244-254)."""
    if layer.original_class.__name__ not in linear_types:
        return None
    deq = layer.sdnq_dequantizer
    if not IS_FP8_MM_SUPPORTED and deq.quantized_matmul_dtype in {"fp8", "float8_e4m3fn"}:
        return False
    if check_param_name_in(param_name, cfg.modules_to_not_use_matmul) is not None:
        return None
    if requested:
        n, k = deq.original_shape
        return bool(min(n, k) >= 32 and n % 16 == 0 and k % 16 == 0)
    return requested


def apply_sdnq_options_to_module(model, quantization_config, dtype=None, dequantize_fp32=None, use_quantized_matmul=None, full_param_name=""):
    """reference loader.py:221-312."""
    if not list(model.children()):
        if dtype is not None and getattr(model, "dtype", torch.float32) not in (torch.float32, torch.float64):
            model = model.to(dtype=dtype)
        return model
    for name, child in list(model.named_children()):
        param_name = f"{full_param_name}.{name}" if full_param_name else name
        if not hasattr(child, "sdnq_dequantizer"):
            setattr(model, name, apply_sdnq_options_to_module(child, quantization_config, dtype=dtype, dequantize_fp32=dequantize_fp32,
                                                              use_quantized_matmul=use_quantized_matmul, full_param_name=param_name))
            continue
        layer, deq = child, child.sdnq_dequantizer
        param_name += ".weight"
        target = _wanted_matmul(layer, param_name, quantization_config, use_quantized_matmul)
        if dtype is not None and deq.result_dtype not in (torch.float32, torch.float64):
            deq.result_dtype = dtype
            if layer.svd_up is not None:
                layer.svd_up.data = layer.svd_up.to(dtype=dtype)
                layer.svd_down.data = layer.svd_down.to(dtype=dtype)
        mm = dtype_dict[deq.quantized_matmul_dtype]
        will_matmul = bool(target) or (target is None and deq.use_quantized_matmul)
        upcast = bool(dequantize_fp32 or dtype_dict[deq.weights_dtype]["num_bits"] > 8
                      or (will_matmul and not mm["is_integer"] and (not USE_TENSORWISE_FP8_MATMUL or mm["num_bits"] == 16)))
        if upcast:
            scale_dtype = layer.scale.dtype if layer.scale.dtype in (torch.float32, torch.float64) else (
                torch.float64 if deq.result_dtype == torch.float64 else torch.float32)
        elif dequantize_fp32 is None and layer.scale.dtype in (torch.float32, torch.float64):
            scale_dtype = layer.scale.dtype
        else:
            scale_dtype = deq.result_dtype
        layer.scale.data = layer.scale.to(dtype=scale_dtype)
        if layer.zero_point is not None:
            layer.zero_point.data = layer.zero_point.to(dtype=scale_dtype)
        if target is not None:
            if target != deq.use_quantized_matmul:
                if not deq.re_quantize_for_matmul and not dtype_dict[deq.weights_dtype]["is_packed"]:
                    layer.scale.data = layer.scale.t_().contiguous()
                    layer.weight.data = layer.weight.t_()
                    if layer.zero_point is not None:
                        layer.zero_point.data = layer.zero_point.t_().contiguous()
                    if target:
                        layer.weight.data = prepare_weight_for_matmul(layer.weight, matmul_dtype=deq.quantized_matmul_dtype)
                    else:
                        layer.scale.data = layer.scale.contiguous()
                        layer.weight.data = layer.weight.contiguous()
                if layer.svd_up is not None:
                    layer.svd_up.data, layer.svd_down.data = prepare_svd_for_matmul(layer.svd_up.t_(), layer.svd_down.t_(), target)
                deq.use_quantized_matmul = target
                layer.forward_func = get_forward_func(layer.original_class.__name__, deq.quantized_matmul_dtype, target)
                layer.__dict__.pop("_sdnq_mm_cache", None)
            wants = use_quantized_matmul or (use_quantized_matmul is None and quantization_config.use_quantized_matmul)
            if not deq.use_quantized_matmul and wants and check_param_name_in(param_name, quantization_config.modules_to_not_use_matmul) is None:
                quantization_config.modules_to_not_use_matmul.append(param_name)
        setattr(model, name, layer)
    return model


def apply_sdnq_options_to_model(model, dtype=None, dequantize_fp32=None, use_quantized_matmul=None):
    """reference loader.py:315-346."""
    model = apply_sdnq_options_to_module(model, model.quantization_config, dtype=dtype, dequantize_fp32=dequantize_fp32,
                                         use_quantized_matmul=use_quantized_matmul)
    holders = [getattr(model, "quantization_config", None)]
    cfg = getattr(model, "config", None)
    if cfg is not None:
        holders.append(getattr(cfg, "quantization_config", None))
        if hasattr(cfg, "get"):
            try:
                holders.append(cfg.get("quantization_config", None))
            except Exception:
                pass
    if hasattr(model, "hf_quantizer"):
        holders.append(getattr(model.hf_quantizer, "quantization_config", None))
    for holder in holders:
        if holder is None or isinstance(holder, dict):
            continue
        if use_quantized_matmul is not None:
            holder.use_quantized_matmul = use_quantized_matmul
        if dequantize_fp32 is not None:
            holder.dequantize_fp32 = dequantize_fp32
    from .siblings import fuse_sibling_projections
    fuse_sibling_projections(model)      # layers may have changed path (matmul on / off): register the sibling groups of the new state
    return model


# ------------------------------------------------------------------------------------------------ checkpoints
def save_sdnq_model(model: torch.nn.Module, model_path: str, max_shard_size: str = "5GB", is_pipeline: bool = False, sdnq_config=None):
    """Write weights as safetensors + `quantization_config.json` (same file names as the reference, loader.py:48-79).
    HF models go through `save_pretrained`; plain nn.Modules get a single `model.safetensors`."""
    os.makedirs(model_path, exist_ok=True)
    cfg = sdnq_config or getattr(model, "quantization_config", None)
    if hasattr(model, "save_pretrained"):
        model.save_pretrained(model_path, max_shard_size=max_shard_size)
    else:
        from safetensors.torch import save_file
        save_file({k: v.contiguous() for k, v in model.state_dict().items()}, os.path.join(model_path, "model.safetensors"))
    if cfg is not None and not is_pipeline:
        d = cfg.to_dict() if hasattr(cfg, "to_dict") else dict(cfg)
        with open(os.path.join(model_path, "quantization_config.json"), "w", encoding="utf-8") as f:
            json.dump(d, f, indent=2, default=str)


def load_sdnq_state_dict(model: torch.nn.Module, model_path: str, device=None, dtype=None, dequantize_fp32=None, use_quantized_matmul=None):
    """Materialise SDNQ modules in `model` from `quantization_config.json`, then assign the safetensors tensors
    (reference loader.py:82-196 for the non-HF branch: config -> sdnq_post_load_quant(pre_quantized=True) -> load_state_dict(assign=True)
    -> post_process_model -> apply options)."""
    from safetensors.torch import load_file
    with open(os.path.join(model_path, "quantization_config.json"), encoding="utf-8") as f:
        raw = json.load(f)
    for key in ("is_integer", "is_unsigned", "quant_method"):
        raw.pop(key, None)
    cfg = SDNQConfig(**raw)
    model = sdnq_post_load_quant(model, quantization_config=cfg, pre_quantized=True, torch_dtype=dtype)
    state = {}
    for fname in sorted(os.listdir(model_path)):
        if fname.endswith(".safetensors"):
            state.update(load_file(os.path.join(model_path, fname), device=str(device) if device is not None else "cpu"))
    model.load_state_dict(state, assign=True)
    model = post_process_model(model)
    if dtype is not None or dequantize_fp32 is not None or use_quantized_matmul is not None:
        model = apply_sdnq_options_to_model(model, dtype=dtype, dequantize_fp32=dequantize_fp32, use_quantized_matmul=use_quantized_matmul)
    return model


def _resolve_model_class(model_config: dict):
    name = model_config.get("_class_name") or model_config.get("architectures")
    if isinstance(name, list):
        name = name[0] if name else None
    if name is None:
        return None
    for package in ("diffusers", "transformers"):
        try:
            module = __import__(package)
        except ImportError:
            continue
        cls = getattr(module, name, None)
        if cls is not None:
            return cls
    return None


def _drop_quantization_config(config):
    if hasattr(config, "quantization_config"):
        try:
            del config.quantization_config
        except Exception:      # noqa: BLE001
            pass
    if hasattr(config, "pop"):
        config.pop("quantization_config", None)
    return config


def load_sdnq_model(model_path: str, model_cls=None, file_name: str | None = None, dtype: torch.dtype | None = None, device="cpu",
                    dequantize_fp32: bool | None = None, use_quantized_matmul: bool | None = None, model_config: dict | None = None,
                    quantization_config=None, load_method: str = "safetensors") -> torch.nn.Module:
    """Build `model_cls` without allocating its weights, swap in SDNQ modules from the stored quantization config, read the
    safetensors shards (file_loader.load_files: 'safetensors' | 'threaded' | 'streamer') and assign them
    (reference loader.py:82-196: same arguments, same order of steps).  The skeleton is created on the meta device, which is
    what accelerate's init_empty_weights does."""
    from .file_loader import load_files
    config_path = os.path.join(model_path, "config.json")
    qconfig_path = os.path.join(model_path, "quantization_config.json")
    if model_config is None:
        model_config = {}
        if os.path.exists(config_path):
            with open(config_path, encoding="utf-8") as f:
                model_config = json.load(f)
    if quantization_config is None:
        if os.path.exists(qconfig_path):
            with open(qconfig_path, encoding="utf-8") as f:
                quantization_config = json.load(f)
        else:
            quantization_config = model_config.get("quantization_config")
            if quantization_config is None:
                raise ValueError(f"Cannot determine quantization_config for {model_path}, please provide quantization_config argument")
    if not isinstance(quantization_config, SDNQConfig):
        raw = dict(quantization_config)
        for key in ("is_integer", "is_unsigned", "quant_method"):
            raw.pop(key, None)
        quantization_config = SDNQConfig(**raw)
    if model_cls is None:
        model_cls = _resolve_model_class(model_config)
    if model_cls is None:
        raise ValueError(f"Cannot determine model class for {model_path}, please provide model_cls argument")
    with torch.device("meta"):
        if hasattr(model_cls, "load_config") and hasattr(model_cls, "from_config"):        # Diffusers
            model = model_cls.from_config(_drop_quantization_config(model_cls.load_config(model_path)))
        elif hasattr(model_cls, "_from_config"):                                             # Transformers
            import transformers
            model = model_cls(_drop_quantization_config(transformers.AutoConfig.from_pretrained(model_path)))
        else:                                                                                # a plain nn.Module taking its config as keywords
            model = model_cls(**_drop_quantization_config(dict(model_config)))
        model = sdnq_post_load_quant(model, torch_dtype=dtype, quantization_config=quantization_config, pre_quantized=True)
    if file_name:
        files = [os.path.join(model_path, file_name)]
    else:
        files = sorted(os.path.join(model_path, f) for f in os.listdir(model_path) if f.endswith(".safetensors"))
    state = load_files(files, key_mapping=getattr(model, "_checkpoint_conversion_mapping", None), device=device, method=load_method)
    tied = getattr(model, "_tied_weights_keys", None)
    if isinstance(tied, dict):
        for key, source in tied.items():
            if source in state and key not in state:
                state[key] = state[source]
    model.load_state_dict(state, assign=True)
    del state
    model.quantization_config = quantization_config
    if hasattr(model, "config"):
        try:
            model.config.quantization_config = quantization_config
        except Exception:      # noqa: BLE001
            pass
    model = post_process_model(model)
    if dtype is not None or dequantize_fp32 is not None or use_quantized_matmul is not None:
        model = apply_sdnq_options_to_model(model, dtype=dtype, dequantize_fp32=dequantize_fp32, use_quantized_matmul=use_quantized_matmul)
    from .siblings import fuse_sibling_projections
    fuse_sibling_projections(model)
    return model
