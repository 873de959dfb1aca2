"""Quantise-time host logic: decide the stored layout of one layer's weight and swap modules in a model.

The layout rules (which tensors exist, their shapes, strides and dtypes) are part of the drop-in boundary because
pre-quantised checkpoints carry no per-layer metadata: it is re-derived by re-running these rules (reference
quantizer.py:66-276, 422-597).  tests/test_host_api.py checks that the same float weight produces bit-identical stored
tensors and metadata as the reference did (fixtures in tests/golden/).

Platform constants folded in (the reference resolves them at import time from the device, kernel_wrappers.py:9-103;
SURVEY.md 5a gives their values on a B200): K-major B operands (use_contiguous_*_mm = False), fp8 matmul supported,
use_tensorwise_fp8_matmul = True."""
import warnings

import torch

from .common import conv_transpose_types, conv_types, dtype_dict, embedding_types, linear_types, weights_dtype_order
from .config import QuantizationMethod, SDNQConfig
from .dequantizer import SDNQDequantizer
from .forward import conv_matmul_unsupported, get_forward_func
from .layers import get_sdnq_wrapper_class
from .packing import pack_float, pack_int
from .quant_math import (apply_hadamard, apply_svdquant, prepare_svd_for_matmul, prepare_weight_for_matmul, quantize_weight,
                         quantize_weight_codebook)
from .utils import (add_module_skip_keys, check_param_name_in, check_quant_is_allowed, check_quantized_matmul_is_allowed,
                    get_quant_kwargs, get_quantized_matmul_dtype)

USE_TENSORWISE_FP8_MATMUL = True
IS_FP8_MM_SUPPORTED = True


def needs_requantize_for_matmul(weights_dtype: str, quantized_matmul_dtype: str, use_codebook: bool) -> bool:
    """Can the stored codes be fed to the matmul as they are?  No if: codebook; wider than the matmul dtype; integer/float
    mismatch; unsigned codes into a float matmul; a packed minifloat that is not a strict subset of the matmul float
    (reference quantizer.py:104-118)."""
    w, m = dtype_dict[weights_dtype], dtype_dict[quantized_matmul_dtype]
    if use_codebook or w["num_bits"] > m["num_bits"] or w["is_integer"] != m["is_integer"]:
        return True
    if w["is_unsigned"] and not m["is_integer"]:
        return True
    packed_float_pair = w["is_packed"] and not w["is_integer"] and not m["is_integer"]
    return bool(packed_float_pair and (w["num_bits"] >= m["num_bits"] or w["max"] > m["max"]))


def _auto_group_size(bits: int, is_linear: bool, has_svd: bool, use_codebook: bool) -> int:
    exponent = 1 + bits + int(is_linear) + int(has_svd) + (3 if use_codebook else 0)
    return 2 ** exponent


def _fit_group_size(group_size: int, channel_size: int):
    """largest group count <= channel_size // group_size that divides the channels (reference quantizer.py:186-200)"""
    if group_size >= channel_size:
        return channel_size, 1
    groups = channel_size // group_size
    while groups * group_size != channel_size:
        groups -= 1
        if groups <= 1:
            return channel_size, 1
        group_size = channel_size // groups
    return int(group_size), int(groups)


@torch.no_grad()
def _quant_kernel_ok(weight, winfo, is_linear, reduction_axes, use_codebook, stochastic) -> bool:
    """K8 covers Linear weights on a CUDA device, integer and float formats of 2..8 bits (intN / uintN, float8_e4m3fn / e5m2, the
    packed eXmY minifloats), scale groups along K that are multiples of 8 (or row-wise), round-to-nearest.  SDNQ_B200_QUANT_KERNEL=0 keeps the eager tensor ops."""
    import os
    if os.environ.get("SDNQ_B200_QUANT_KERNEL", "1") in ("0", "false", "no") or not weight.is_cuda or not is_linear or use_codebook or stochastic:
        return False
    if not 2 <= winfo["num_bits"] <= 8 or reduction_axes != -1 or weight.ndim not in (2, 3):
        return False
    if not (winfo["is_integer"] or winfo["is_packed"] or winfo["torch_dtype"] in (torch.float8_e4m3fn, torch.float8_e5m2)):
        return False
    if weight.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        return False
    return weight.shape[-1] % 8 == 0


def sdnq_quantize_layer_weight(
    weight: torch.Tensor,
    layer_class_name: str | None = None,
    weights_dtype: str = "int8",
    quantized_matmul_dtype: str | None = None,
    group_size: int = 0,
    hadamard_group_size: int = 256,
    svd_rank: int = 32,
    svd_steps: int = 8,
    codebook_steps: int = 24,
    use_svd: bool = False,
    use_hadamard: bool = False,
    use_codebook: bool = False,
    use_quantized_matmul: bool = False,
    use_stochastic_rounding: bool = False,
    dequantize_fp32: bool = True,
    hadamard: torch.Tensor | None = None,
    using_pre_calculated_svd: bool = False,
    using_pre_rotated_hadamard: bool = False,
    skip_sr: bool = False,
    param_name: str | None = None,
    torch_dtype: torch.dtype | None = None,
):
    """float weight -> (SDNQDequantizer, {"weight", "scale", "zero_point", "svd_up", "svd_down"})."""
    weight = weight.detach()
    original_shape, original_stride = weight.shape, weight.stride()
    torch_dtype = torch_dtype or weight.dtype
    quantized_matmul_dtype = get_quantized_matmul_dtype(weights_dtype, quantized_matmul_dtype)
    winfo, minfo = dtype_dict[weights_dtype], dtype_dict[quantized_matmul_dtype]
    requant = needs_requantize_for_matmul(weights_dtype, quantized_matmul_dtype, use_codebook)

    is_conv = layer_class_name in conv_types
    is_conv_t = layer_class_name in conv_transpose_types
    is_linear = layer_class_name in linear_types
    result_shape = None
    if is_conv:
        reduction_axes = 1
        out_ch, in_ch = weight.shape[:2]
        use_quantized_matmul = check_quantized_matmul_is_allowed(use_quantized_matmul, out_ch, in_ch)
        if use_quantized_matmul and not requant and not winfo["is_packed"]:
            result_shape = weight.shape
            weight = weight.flatten(1, -1)
            reduction_axes = -1
    elif is_conv_t:
        reduction_axes = 0
        in_ch, out_ch = weight.shape[:2]
        use_quantized_matmul = False
    elif is_linear:
        reduction_axes = -1
        out_ch, in_ch = weight.shape
        use_quantized_matmul = check_quantized_matmul_is_allowed(use_quantized_matmul, out_ch, in_ch)
    else:
        out_ch, in_ch = (weight.shape[-2:] if weight.ndim > 1 else (1, weight.shape[-1]))
        reduction_axes = -1
        use_quantized_matmul = False

    # scales stay fp32 unless the user opted out and the format tolerates it (reference quantizer.py:147-156)
    scale_dtype = None
    float_mm_needs_fp32 = use_quantized_matmul and not minfo["is_integer"] and (not USE_TENSORWISE_FP8_MATMUL or minfo["num_bits"] == 16)
    if not dequantize_fp32 and winfo["max"] <= 16384 and not float_mm_needs_fp32:
        scale_dtype = torch_dtype

    if use_hadamard:
        weight, use_hadamard, hadamard_group_size = apply_hadamard(weight, group_size=hadamard_group_size, hadamard=hadamard, layer_class_name=layer_class_name)

    svd_up = svd_down = None
    if use_svd:
        try:
            weight, svd_up, svd_down = apply_svdquant(weight, rank=svd_rank, steps=svd_steps, dtype=torch_dtype)
            if use_quantized_matmul:
                svd_up, svd_down = svd_up.t_(), svd_down.t_()
            svd_up, svd_down = prepare_svd_for_matmul(svd_up, svd_down, use_quantized_matmul)
        except Exception:  # the reference silently drops SVD when the decomposition fails (quantizer.py:162-169)
            svd_up = svd_down = None

    if group_size == 0:
        if use_quantized_matmul and not requant and winfo["num_bits"] >= 6:
            group_size = -1
        else:
            group_size = _auto_group_size(winfo["num_bits"], is_linear, svd_up is not None or using_pre_calculated_svd, use_codebook)

    num_groups = 1
    if group_size > 0:
        group_size, num_groups = _fit_group_size(group_size, in_ch)
        if num_groups > 1:
            if result_shape is None:
                result_shape = weight.shape
            if is_conv:
                reduction_axes = 2
                weight = weight.unflatten(1, (num_groups, group_size))
            elif is_conv_t:
                reduction_axes = 1
                weight = weight.unflatten(1, (group_size, num_groups))
            else:
                reduction_axes = -1
                weight = weight.unflatten(-1, (num_groups, group_size))
        else:
            group_size = -1
    elif group_size == -2:
        reduction_axes = tuple(range(weight.ndim))

    requant = requant or num_groups > 1
    transpose_weights = bool(use_quantized_matmul and not requant and not winfo["is_packed"])
    cast_scale = not (transpose_weights and not USE_TENSORWISE_FP8_MATMUL and not minfo["is_integer"])
    cast_to = scale_dtype if cast_scale else None

    packed_by_kernel = False
    if _quant_kernel_ok(weight, winfo, is_linear, reduction_axes, use_codebook, use_stochastic_rounding and not skip_sr):
        # K8: scale + round + clamp + pack in one pass over the weight on the GPU (csrc/weight_quant.cu) -- the arithmetic of
        # quantize_weight + pack_int / pack_float below, bit for bit
        from . import ops
        view_shape = weight.shape
        codes, scale, zero_point = ops.quantize_weight(weight.reshape(out_ch, -1), weights_dtype, group_size if num_groups > 1 else -1, cast_to)
        scale = scale.view(*view_shape[:-1], 1)
        zero_point = None if zero_point is None else zero_point.view(*view_shape[:-1], 1)
        if cast_to is not None:
            scale = scale.to(cast_to)
            zero_point = None if zero_point is None else zero_point.to(cast_to)
        if winfo["is_packed"]:
            words = {2: 1, 4: 1, 3: 3, 5: 5, 6: 3, 7: 7, 8: 0}[winfo["num_bits"]]
            if words == 0:                       # 8-bit eXmY minifloats: one uint8 code per weight, in the weight's own shape (pack_float)
                packed_weight, packed_by_kernel = codes.view(view_shape), True
            else:
                packed_weight, packed_by_kernel = (codes if words == 1 else codes.view(-1, words)), True
            weight = torch.empty(view_shape, dtype=torch.uint8, device="meta")       # only its shape is used below
        else:
            weight = codes.view(view_shape)
    elif use_codebook:
        weight, scale = quantize_weight_codebook(weight, reduction_axes, weights_dtype, dtype=cast_to, steps=codebook_steps)
        zero_point = None
    else:
        weight, scale, zero_point = quantize_weight(weight, reduction_axes, weights_dtype, dtype=cast_to,
                                                    use_stochastic_rounding=(use_stochastic_rounding and not skip_sr))

    if transpose_weights:          # store the matmul operand: logical [K,N], K-major in memory; scale / zp as [1,N]
        scale = scale.t_().contiguous()
        weight = weight.t_()
        if zero_point is not None:
            zero_point = zero_point.t_().contiguous()
        weight = prepare_weight_for_matmul(weight, matmul_dtype=quantized_matmul_dtype)

    quantized_weight_shape = weight.shape
    if packed_by_kernel:
        weight = packed_weight
    elif winfo["is_packed"]:
        weight = pack_int(weight, weights_dtype) if winfo["is_integer"] else pack_float(weight, weights_dtype)
    else:
        weight = weight.to(dtype=winfo["torch_dtype"])

    dequantizer = SDNQDequantizer(
        result_dtype=torch_dtype, result_shape=result_shape, original_shape=original_shape, original_stride=original_stride,
        quantized_weight_shape=quantized_weight_shape, weights_dtype=weights_dtype, quantized_matmul_dtype=quantized_matmul_dtype,
        hadamard_group_size=hadamard_group_size, group_size=group_size, svd_rank=svd_rank, svd_steps=svd_steps,
        codebook_steps=codebook_steps, use_quantized_matmul=use_quantized_matmul, re_quantize_for_matmul=requant,
        use_stochastic_rounding=use_stochastic_rounding, use_hadamard=bool(use_hadamard or using_pre_rotated_hadamard),
        use_codebook=use_codebook, layer_class_name=layer_class_name)
    return dequantizer, {"weight": weight, "scale": scale, "zero_point": zero_point, "svd_up": svd_up, "svd_down": svd_down}


@torch.no_grad()
def sdnq_quantize_layer_weight_dynamic(weight, layer_class_name=None, weights_dtype="uint4", quantized_matmul_dtype=None, group_size=0,
                                       hadamard_group_size=256, svd_rank=32, svd_steps=8, codebook_steps=24, dynamic_loss_threshold=None,
                                       use_svd=False, use_hadamard=False, use_codebook=False, use_quantized_matmul=False,
                                       use_stochastic_rounding=False, dequantize_fp32=True, hadamard=None, param_name=None,
                                       torch_dtype=None, quantization_config=None):
    """Walk `weights_dtype_order` upwards from `weights_dtype` until the normalised MSE of quantise->dequantise drops below the
    threshold (default 10^-(bits/2)); record the choice in the config (reference quantizer.py:280-419).  Needs a CUDA weight:
    the trial dequantisation runs through the K3 kernel."""
    torch_dtype = torch_dtype or weight.dtype
    if dynamic_loss_threshold is None or dynamic_loss_threshold < 0:
        dynamic_loss_threshold = 10 ** -(dtype_dict[weights_dtype]["num_bits"] / 2)
    weight = weight.detach()
    if weight.dtype != torch.float64:
        weight = weight.to(torch.float32, copy=False)
    variance = weight.std().square_().clamp_(min=1e-8)
    original = weight
    if use_hadamard:
        weight, use_hadamard, hadamard_group_size = apply_hadamard(weight, group_size=hadamard_group_size, hadamard=hadamard, layer_class_name=layer_class_name)
    svd = {False: (None, None), True: (None, None)}
    if use_svd:
        try:
            weight, up, down = apply_svdquant(weight, rank=svd_rank, steps=svd_steps, dtype=torch_dtype)
            svd[False] = prepare_svd_for_matmul(up, down, False)
            if use_quantized_matmul:
                svd[True] = prepare_svd_for_matmul(svd[False][0].clone().t_(), svd[False][1].clone().t_(), True)
        except Exception:
            pass

    def done(result, skip_matmul_name=False):
        if quantization_config is None:
            return result
        if result is None:
            quantization_config.modules_to_not_convert.append(param_name)
            return None, quantization_config
        deq = result[0]
        quantization_config.modules_dtype_dict.setdefault(deq.weights_dtype, []).append(param_name)
        if skip_matmul_name and check_param_name_in(param_name, quantization_config.modules_to_not_use_matmul) is None:
            quantization_config.modules_to_not_use_matmul.append(param_name)
        return result, quantization_config

    base_bits = dtype_dict[weights_dtype]["num_bits"]
    for candidate in weights_dtype_order[weights_dtype_order.index(weights_dtype):]:
        cinfo = dtype_dict[candidate]
        if use_codebook and not (cinfo["is_unsigned"] and cinfo["is_integer"]):
            continue
        mm_dtype = get_quantized_matmul_dtype(candidate, quantized_matmul_dtype)
        minfo = dtype_dict[mm_dtype]
        use_mm, flag_no_mm = use_quantized_matmul, False
        if (
            (minfo["is_integer"] and not cinfo["is_integer"])
            or (cinfo["num_bits"] == minfo["num_bits"] and cinfo["is_unsigned"] and not minfo["is_integer"])
            or (base_bits <= minfo["num_bits"] < cinfo["num_bits"])
        ):
            use_mm, flag_no_mm = False, True
        deq, data = sdnq_quantize_layer_weight(
            weight, layer_class_name=layer_class_name, weights_dtype=candidate, quantized_matmul_dtype=mm_dtype, torch_dtype=torch_dtype,
            hadamard_group_size=hadamard_group_size, group_size=group_size, svd_rank=svd_rank, svd_steps=svd_steps, codebook_steps=codebook_steps,
            use_svd=False, use_hadamard=False, use_codebook=use_codebook, use_quantized_matmul=use_mm,
            use_stochastic_rounding=use_stochastic_rounding, dequantize_fp32=dequantize_fp32, using_pre_calculated_svd=use_svd,
            using_pre_rotated_hadamard=use_hadamard, param_name=param_name)
        data["svd_up"], data["svd_down"] = svd[bool(deq.use_quantized_matmul)]
        restored = deq(data["weight"], data["scale"], zero_point=data["zero_point"], svd_up=data["svd_up"], svd_down=data["svd_down"],
                       skip_quantized_matmul=deq.use_quantized_matmul, dtype=weight.dtype)
        loss = torch.nn.functional.mse_loss(original, restored.view_as(original)).div_(variance)
        if loss <= dynamic_loss_threshold:
            return done((deq, data), skip_matmul_name=flag_no_mm)
    return done(None)


@torch.no_grad()
def sdnq_quantize_layer(layer: torch.nn.Module, quantization_config: SDNQConfig, torch_dtype: torch.dtype | None = None,
                        param_name: str = "", quant_kwargs: dict | None = None, pre_quantized: bool = False):
    """Quantise one module in place and return (SDNQ wrapper, config)   (reference quantizer.py:422-473)."""
    torch_dtype = torch_dtype or layer.weight.dtype
    if quant_kwargs is None:
        quant_kwargs = get_quant_kwargs(layer, quantization_config, torch_dtype=torch_dtype, param_name=param_name)
    layer_class_name = layer.__class__.__name__
    wanted_matmul = bool(quant_kwargs["use_quantized_matmul"])
    if wanted_matmul and layer_class_name in conv_types and not pre_quantized:
        # convolutions whose quantized matmul has no kernel here (groups, Conv3d, string padding) stay on the dequant path -- decided
        # now, so that a model never quantises / loads fine and then fails in its first forward
        why = conv_matmul_unsupported(layer)
        if why is not None:
            warnings.warn(f"sdnq_b200: {param_name or layer_class_name}: {why} has no W8A8 conv kernel; keeping the layer on the "
                          "dequant path (use_quantized_matmul=False for it)", stacklevel=2)
            quant_kwargs["use_quantized_matmul"] = False
    is_conv = layer_class_name in conv_types or layer_class_name in conv_transpose_types
    if (layer_class_name in embedding_types and not quantization_config.quant_embedding) or (is_conv and not quantization_config.quant_conv):
        quantization_config.modules_to_not_convert.append(param_name)
        return layer, quantization_config

    return_device = quant_kwargs.pop("return_device")
    quantization_device = quant_kwargs.pop("quantization_device")
    non_blocking = quant_kwargs.pop("non_blocking")
    dynamic = quant_kwargs.pop("use_dynamic_quantization")
    layer.weight.requires_grad_(False)
    if return_device is None:
        return_device = layer.weight.device
    if quantization_device is not None:
        layer.weight.data = layer.weight.to(quantization_device, non_blocking=non_blocking, copy=False)

    if dynamic:
        result, quantization_config = sdnq_quantize_layer_weight_dynamic(layer.weight, quantization_config=quantization_config, **quant_kwargs)
    else:
        result = sdnq_quantize_layer_weight(layer.weight, **quant_kwargs)

    if result is None:
        layer.weight = torch.nn.Parameter(layer.weight.to(return_device, dtype=torch_dtype, non_blocking=non_blocking, copy=False), requires_grad=False)
        return layer, quantization_config

    layer.sdnq_dequantizer, tensors = result
    deq = layer.sdnq_dequantizer
    layer = get_sdnq_wrapper_class(layer, get_forward_func(layer_class_name, deq.quantized_matmul_dtype, deq.use_quantized_matmul))
    for key, value in tensors.items():
        if isinstance(value, torch.Tensor):
            param = torch.nn.Parameter(value.to(return_device, non_blocking=non_blocking, copy=False), requires_grad=False)
            param._is_hf_initialized = True
            setattr(layer, key, param)
        else:
            setattr(layer, key, value)
    if (wanted_matmul and not deq.use_quantized_matmul
            and check_param_name_in(param_name, quantization_config.modules_to_not_use_matmul) is None):
        quantization_config.modules_to_not_use_matmul.append(param_name)
    return layer, quantization_config


@torch.no_grad()
def apply_sdnq_to_module(model: torch.nn.Module, quantization_config: SDNQConfig, torch_dtype=None, pre_quantized: bool = False,
                         full_param_name: str = ""):
    """Depth-first module swap (reference quantizer.py:476-495)."""
    for child_name, child in list(model.named_children()):
        name = f"{full_param_name}.{child_name}" if full_param_name else child_name
        if getattr(child, "weight", None) is not None:
            name = name + ".weight"
            if check_param_name_in(name, quantization_config.modules_to_not_convert) is None:
                if check_quant_is_allowed(child.__class__.__name__, child.weight, quantization_config, pre_quantized=pre_quantized):
                    child, quantization_config = sdnq_quantize_layer(child, quantization_config, torch_dtype=torch_dtype, param_name=name,
                                                                     pre_quantized=pre_quantized)
                    setattr(model, child_name, child)
                else:
                    quantization_config.modules_to_not_convert.append(name)
        child, quantization_config = apply_sdnq_to_module(child, quantization_config, torch_dtype=torch_dtype, pre_quantized=pre_quantized,
                                                          full_param_name=name)
        setattr(model, child_name, child)
    return model, quantization_config


_CONFIG_KEYS = ("weights_dtype", "quantized_matmul_dtype", "hadamard_group_size", "group_size", "svd_rank", "svd_steps", "codebook_steps",
                "dynamic_loss_threshold", "use_svd", "use_hadamard", "use_codebook", "quant_conv", "quant_embedding", "use_quantized_matmul",
                "use_quantized_matmul_conv", "use_dynamic_quantization", "use_stochastic_rounding", "dequantize_fp32", "non_blocking",
                "add_skip_keys", "minimum_allowed_numel", "minimum_allowed_channel_size", "modules_to_not_convert",
                "modules_to_not_use_matmul", "modules_dtype_dict", "modules_quant_config", "quantization_device", "return_device")


@torch.no_grad()
def sdnq_post_load_quant(model: torch.nn.Module, *args, torch_dtype: torch.dtype | None = None, quantization_config: SDNQConfig | None = None,
                         pre_quantized: bool = False, **kwargs) -> torch.nn.Module:
    """Quantise an already-loaded model in place.  Keyword arguments are the SDNQConfig fields (reference quantizer.py:498-597)."""
    if pre_quantized:
        kwargs["add_skip_keys"] = False
        kwargs["use_dynamic_quantization"] = False
        if quantization_config is not None:
            quantization_config.add_skip_keys = False
            quantization_config.use_dynamic_quantization = False
    else:
        cfg_holder = getattr(model, "config", None)
        already = hasattr(model, "quantization_config") or (cfg_holder is not None and (
            hasattr(cfg_holder, "quantization_config") or (isinstance(cfg_holder, dict) and "quantization_config" in cfg_holder)))
        if already:
            raise RuntimeError("Quantizing a pre-quantized model is not supported!")
    if quantization_config is None:
        quantization_config = SDNQConfig(**{k: v for k, v in kwargs.items() if k in _CONFIG_KEYS})
    if quantization_config.add_skip_keys:
        model, quantization_config = add_module_skip_keys(model, quantization_config)
    model.eval()
    model, quantization_config = apply_sdnq_to_module(model, quantization_config, torch_dtype=torch_dtype, pre_quantized=pre_quantized)
    model.quantization_config = quantization_config
    if hasattr(model, "config"):
        try:
            model.config.quantization_config = quantization_config
        except Exception:
            pass
        try:
            model.config["quantization_config"] = quantization_config.to_dict()
        except Exception:
            pass
    model.quantization_method = QuantizationMethod.SDNQ
    from .siblings import fuse_sibling_projections
    fuse_sibling_projections(model)          # to_q / to_k / to_v ...: registered now, their shared operand is built at the first forward
    return model
