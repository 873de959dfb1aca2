"""SDNQ wrapper modules.  Class names, attribute names (`weight, scale, zero_point, svd_up, svd_down, sdnq_dequantizer,
original_class, forward_func`) and `.dequantize()` are the upper drop-in boundary (reference layers/__init__.py:6-93):
loaders, `apply_sdnq_options_to_model` and serialised checkpoints key on them."""
from collections.abc import Callable

import torch

_SKIP_ON_ADOPT = {"forward", "forward_func", "original_class", "state_dict", "load_state_dict"}


class SDNQLayer(torch.nn.Module):
    def __init__(self, original_layer: torch.nn.Module, forward_func: Callable):
        torch.nn.Module.__init__(self)
        # adopt the wrapped module's state wholesale (parameters, buffers, hooks, config attrs); like the reference this shares
        # -- and therefore mutates -- the original module's parameter dict
        for key, value in original_layer.__dict__.items():
            if key not in _SKIP_ON_ADOPT:
                setattr(self, key, value)
        self.original_class = original_layer.__class__
        self.forward_func = forward_func

    @property
    def dtype(self) -> torch.dtype:
        deq = getattr(self, "sdnq_dequantizer", None)
        return deq.result_dtype if deq is not None else self.weight.dtype

    def dequantize(self):
        """Turn this module back into its original class with a dense weight."""
        deq = getattr(self, "sdnq_dequantizer", None)
        if deq is not None:
            dense = deq(self.weight, self.scale, zero_point=self.zero_point, svd_up=self.svd_up, svd_down=self.svd_down,
                        skip_quantized_matmul=deq.use_quantized_matmul)
            self.weight = torch.nn.Parameter(dense, requires_grad=True)
            del self.sdnq_dequantizer, self.scale, self.zero_point, self.svd_up, self.svd_down
            self.__dict__.pop("_sdnq_mm_cache", None)
        self.__class__ = self.original_class
        del self.original_class, self.forward_func
        return self

    def forward(self, *args, **kwargs) -> torch.Tensor:
        return self.forward_func(self, *args, **kwargs)

    # Everything the forwards derive from the stored tensors (the cached matmul operand, the side stream's licence to run ahead
    # of the caller's stream) is tied to a generation counter that any wholesale replacement of the tensors bumps.
    def _sdnq_touch(self):
        self.__dict__["_sdnq_generation"] = self.__dict__.get("_sdnq_generation", 0) + 1

    def _apply(self, fn, *args, **kwargs):
        self._sdnq_touch()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self._sdnq_touch()
        return super()._load_from_state_dict(*args, **kwargs)

    def __getstate__(self):
        # derived per-process state (cached operands, stream bookkeeping: weak references, device buffers) is not part of the module
        return {k: v for k, v in self.__dict__.items() if not k.startswith("_sdnq_")}

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(original_class={self.original_class} forward_func={self.forward_func} "
                f"sdnq_dequantizer={getattr(self, 'sdnq_dequantizer', None)})")


class SDNQLinear(SDNQLayer, torch.nn.Linear):
    original_class: torch.nn.Linear


class SDNQEmbedding(SDNQLayer, torch.nn.Embedding):
    original_class: torch.nn.Embedding


class SDNQConv1d(SDNQLayer, torch.nn.Conv1d):
    original_class: torch.nn.Conv1d


class SDNQConv2d(SDNQLayer, torch.nn.Conv2d):
    original_class: torch.nn.Conv2d


class SDNQConv3d(SDNQLayer, torch.nn.Conv3d):
    original_class: torch.nn.Conv3d


class SDNQConvTranspose1d(SDNQLayer, torch.nn.ConvTranspose1d):
    original_class: torch.nn.ConvTranspose1d


class SDNQConvTranspose2d(SDNQLayer, torch.nn.ConvTranspose2d):
    original_class: torch.nn.ConvTranspose2d


class SDNQConvTranspose3d(SDNQLayer, torch.nn.ConvTranspose3d):
    original_class: torch.nn.ConvTranspose3d


_WRAPPERS = {"Linear": SDNQLinear, "Embedding": SDNQEmbedding, "Gemma4TextScaledWordEmbedding": SDNQEmbedding, "Conv1d": SDNQConv1d,
             "Conv2d": SDNQConv2d, "Conv3d": SDNQConv3d, "ConvTranspose1d": SDNQConvTranspose1d, "ConvTranspose2d": SDNQConvTranspose2d,
             "ConvTranspose3d": SDNQConvTranspose3d}
torch.serialization.add_safe_globals([SDNQLayer, *set(_WRAPPERS.values())])


def get_sdnq_wrapper_class(original_layer: torch.nn.Module, forward_func: Callable) -> SDNQLayer:
    return _WRAPPERS.get(original_layer.__class__.__name__, SDNQLayer)(original_layer, forward_func)
