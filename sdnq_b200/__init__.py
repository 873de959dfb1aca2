"""sdnq_b200: B200-native (sm_100a) implementation of the SDNQ quantized-Linear (and Conv) forward.

Public surface mirrors the reference package (`sdnq/__init__.py:1-16`): configure with `SDNQConfig`, quantise a loaded model
with `sdnq_post_load_quant`, flip run-time options with `apply_sdnq_options_to_model`.  The forward of every `SDNQLinear`
runs hand-written CUDA kernels from `libsdnq_b200.so` through the C ABI in `include/sdnq_b200.h`."""
from .attention import sdnq_attention, sdnq_triton_atten  # noqa: F401
from .common import dtype_dict, sdnq_version  # noqa: F401
from .config import QuantizationMethod, SDNQConfig  # noqa: F401
from .dequantizer import SDNQDequantizer  # noqa: F401
from .forward import get_forward_func  # noqa: F401
from .layers import (SDNQConv1d, SDNQConv2d, SDNQConv3d, SDNQConvTranspose1d, SDNQConvTranspose2d, SDNQConvTranspose3d,  # noqa: F401
                     SDNQLayer, SDNQLinear)
from .loader import (apply_sdnq_options_to_model, load_sdnq_model, load_sdnq_state_dict, post_process_model, save_sdnq_model)  # noqa: F401
from .siblings import fuse_named_siblings, fuse_sibling_projections, group_siblings  # noqa: F401
from .quantizer import (apply_sdnq_to_module, sdnq_post_load_quant, sdnq_quantize_layer,  # noqa: F401
                        sdnq_quantize_layer_weight)

__version__ = "0.1.0"
__all__ = ["apply_sdnq_options_to_model", "load_sdnq_model", "load_sdnq_state_dict", "post_process_model", "save_sdnq_model", "SDNQConfig", "SDNQDequantizer", "SDNQLayer", "SDNQLinear", "SDNQConv1d", "SDNQConv2d", "SDNQConv3d", "SDNQConvTranspose1d", "SDNQConvTranspose2d", "SDNQConvTranspose3d", "QuantizationMethod", "apply_sdnq_to_module", "dtype_dict",
           "fuse_named_siblings", "fuse_sibling_projections", "group_siblings", "get_forward_func", "sdnq_attention", "sdnq_post_load_quant", "sdnq_quantize_layer", "sdnq_quantize_layer_weight", "sdnq_triton_atten", "sdnq_version"]
