"""Development-time loader for the upstream reference (container only).

Used ONLY by tests/golden/generate.py to produce the committed fixtures.
Nothing under tests/ that runs on the GPU box imports this file:
/root/reference does not exist there.

Recipe: SURVEY.md Appendix B.  The reference imports diffusers/accelerate at
module top (quantizer.py:7-12) and neither is installed, so tiny stubs are put
in sys.modules first.
"""
import contextlib
import os
import sys
import types

REF_SRC = "/root/reference/src"


def load_reference():
    if "sdnq" in sys.modules:
        return sys.modules["sdnq"]
    if not os.path.isdir(REF_SRC):
        raise RuntimeError("reference sources not present (only available in the authoring container)")
    os.environ.setdefault("SDNQ_USE_TORCH_COMPILE", "0")
    import transformers.quantizers  # noqa: F401  (must precede the accelerate stub)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class DiffusersQuantizer:
        def __init__(self, quantization_config=None, **kw):
            self.quantization_config = quantization_config
            self.pre_quantized = kw.get("pre_quantized", False)

    class QuantizationConfigMixin:
        @classmethod
        def from_dict(cls, d, **kw):
            return cls(**d)

    def get_module_from_name(module, name):
        parts = name.split(".")
        for p in parts[:-1]:
            module = getattr(module, p)
        return module, parts[-1]

    mod("diffusers", __version__="0.40.0", __path__=[])
    mod("diffusers.quantizers", __path__=[])
    mod("diffusers.quantizers.base", DiffusersQuantizer=DiffusersQuantizer)
    mod("diffusers.quantizers.quantization_config", QuantizationConfigMixin=QuantizationConfigMixin)
    mod("diffusers.quantizers.auto", AUTO_QUANTIZER_MAPPING={}, AUTO_QUANTIZATION_CONFIG_MAPPING={})
    mod("diffusers.utils", get_module_from_name=get_module_from_name)
    mod("accelerate", init_empty_weights=contextlib.nullcontext)
    sys.path.insert(0, REF_SRC)
    import sdnq
    return sdnq
