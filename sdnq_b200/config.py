"""SDNQConfig: the user-facing quantisation config (same fields, defaults, validation and JSON form as the
reference's, quantizer.py:845-1082, so `quantization_config.json` files of pre-quantised repos load unchanged)."""
from enum import Enum

import torch

from .common import accepted_matmul_dtypes, accepted_weight_dtypes, dtype_dict, sdnq_version as _current_version

try:  # makes SDNQConfig a first-class HF quantization config when transformers is around; not needed otherwise
    from transformers.utils.quantization_config import QuantizationConfigMixin as _ConfigBase
except Exception:  # pragma: no cover
    class _ConfigBase:  # type: ignore
        @classmethod
        def from_dict(cls, config_dict, return_unused_kwargs=False, **kwargs):
            cfg = cls(**config_dict)
            return (cfg, kwargs) if return_unused_kwargs else cfg


class QuantizationMethod(str, Enum):
    SDNQ = "sdnq"
    SDNQ_TRAINING = "sdnq_training"


def _as_list(value, what):
    if value is None:
        return []
    if isinstance(value, str):
        return [value]
    if isinstance(value, (tuple, list)):
        return list(value)
    raise ValueError(f"{what} must be a list but got {type(value)}")


class SDNQConfig(_ConfigBase):
    """See the reference docstring (quantizer.py:846-936) for the meaning of every option; the semantics are identical.

    weights_dtype            storage dtype of the weights ("int8", "uint4", "float8_e4m3fn", "float6_e3m2fn", ...)
    quantized_matmul_dtype   None -> "int8" for integer weights ("uint8" for uint8), "float8_e4m3fn" for <16-bit floats
    group_size               0 auto, -1 row-wise, -2 tensor-wise, >0 elements per scale along the input channels
    use_quantized_matmul     run the Linear as W8A8 on the int8 / fp8 tensor cores instead of dequant + bf16 GEMM
    use_svd / svd_rank       SVDQuant low-rank correction;  use_hadamard / hadamard_group_size   Hadamard rotation
    """

    def __init__(
        self,
        weights_dtype: str = "int8",
        quantized_matmul_dtype: str | None = None,
        hadamard_group_size: int = 256,
        group_size: int = 0,
        svd_rank: int = 32,
        svd_steps: int = 8,
        codebook_steps: int = 24,
        dynamic_loss_threshold: float | None = None,
        use_svd: bool = False,
        use_hadamard: bool = False,
        use_codebook: bool = False,
        use_grad_ckpt: bool = True,
        quant_conv: bool = False,
        quant_embedding: bool = False,
        use_quantized_matmul: bool = False,
        use_quantized_matmul_conv: bool = False,
        use_static_quantization: bool = True,
        use_dynamic_quantization: bool = False,
        use_stochastic_rounding: bool = False,
        dequantize_fp32: bool = True,
        non_blocking: bool = False,
        add_skip_keys: bool = True,
        minimum_allowed_numel: int = 16384,
        minimum_allowed_channel_size: int = 32,
        modules_to_not_convert: list[str] | None = None,
        modules_to_not_use_matmul: list[str] | None = None,
        modules_dtype_dict: dict[str, list[str]] | None = None,
        modules_quant_config: dict[str, dict] | None = None,
        quantization_device: torch.device | None = None,
        return_device: torch.device | None = None,
        sdnq_version: str | None = None,
        is_training: bool = False,
        **kwargs,
    ):
        self.weights_dtype = weights_dtype
        self.quantized_matmul_dtype = quantized_matmul_dtype
        self.is_training = is_training
        self.hadamard_group_size = hadamard_group_size
        self.group_size = group_size
        self.svd_rank = svd_rank
        self.dynamic_loss_threshold = dynamic_loss_threshold
        self.use_svd = use_svd
        self.svd_steps = svd_steps
        self.codebook_steps = codebook_steps
        self.use_hadamard = use_hadamard
        self.use_codebook = use_codebook
        self.use_grad_ckpt = use_grad_ckpt
        self.quant_conv = quant_conv
        self.quant_embedding = quant_embedding
        self.use_quantized_matmul = use_quantized_matmul
        self.use_quantized_matmul_conv = use_quantized_matmul_conv
        self.use_static_quantization = use_static_quantization
        self.use_dynamic_quantization = use_dynamic_quantization
        self.use_stochastic_rounding = use_stochastic_rounding
        self.dequantize_fp32 = dequantize_fp32
        self.non_blocking = non_blocking
        self.add_skip_keys = add_skip_keys
        self.minimum_allowed_numel = minimum_allowed_numel
        self.minimum_allowed_channel_size = minimum_allowed_channel_size
        self.modules_to_not_convert = modules_to_not_convert
        self.modules_to_not_use_matmul = modules_to_not_use_matmul
        self.modules_dtype_dict = modules_dtype_dict
        self.modules_quant_config = modules_quant_config
        self.quantization_device = quantization_device
        self.return_device = return_device
        self.sdnq_version = _current_version if sdnq_version is None else sdnq_version
        if weights_dtype not in accepted_weight_dtypes:
            raise ValueError(f"SDNQ only support weight dtypes in {sorted(accepted_weight_dtypes)} but found {weights_dtype}")
        self.is_integer = dtype_dict[weights_dtype]["is_integer"]
        self.is_unsigned = dtype_dict[weights_dtype]["is_unsigned"]
        self.quant_method = QuantizationMethod.SDNQ_TRAINING if is_training else QuantizationMethod.SDNQ
        self.post_init()

    def post_init(self) -> None:
        if self.quantized_matmul_dtype is not None and self.quantized_matmul_dtype not in accepted_matmul_dtypes:
            raise ValueError(f"SDNQ only support quantized matmul dtypes in {accepted_matmul_dtypes} but found {self.quantized_matmul_dtype}")
        if self.use_codebook and not (self.is_integer and self.is_unsigned):
            raise ValueError(f"SDNQ: use_codebook is only supported with unsigned integer dtypes but found {self.weights_dtype}")
        self.modules_to_not_convert = list(set(_as_list(self.modules_to_not_convert, "modules_to_not_convert")))
        self.modules_to_not_use_matmul = list(set(_as_list(self.modules_to_not_use_matmul, "modules_to_not_use_matmul")))
        if self.modules_dtype_dict is None:
            self.modules_dtype_dict = {}
        elif not isinstance(self.modules_dtype_dict, dict):
            raise ValueError(f"modules_dtype_dict must be a dict but got {type(self.modules_dtype_dict)}")
        normalised = {}
        for key, names in self.modules_dtype_dict.items():
            if isinstance(names, str):
                names = [names]
            elif isinstance(names, tuple):
                names = list(names)
            if not isinstance(key, str) or not isinstance(names, list):
                raise TypeError(f"modules_dtype_dict must be a dictionary of strings and lists but got {type(key)} and {type(names)}")
            normalised[key] = list(set(names))
        self.modules_dtype_dict = normalised
        self.modules_quant_config = dict(self.modules_quant_config) if self.modules_quant_config is not None else {}

    def to_dict(self) -> dict:
        out = self.__dict__.copy()
        for key in ("quantization_device", "return_device"):
            out[key] = str(out[key]) if out[key] is not None else None
        return out

    def __repr__(self) -> str:
        shown = ("weights_dtype", "group_size", "use_quantized_matmul", "quantized_matmul_dtype", "use_hadamard", "hadamard_group_size",
                 "use_svd", "svd_rank", "use_codebook", "dequantize_fp32", "quant_conv", "quant_embedding")
        return "SDNQConfig(" + " ".join(f"{k}={getattr(self, k)}" for k in shown) + ")"

    __str__ = __repr__
