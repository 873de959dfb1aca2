"""Import the UNMODIFIED upstream reference (Disty0/sdnq) as the module `sdnq`  --  TEST / BENCH INFRASTRUCTURE ONLY.

Search order: `oracle/_ref/` (a pristine copy of the reference's `src/sdnq` tree made by `oracle/build_ref.py`; git-ignored,
travels to the GPU box with the gpurun snapshot), then `/root/reference/src` (authoring container only).  Nothing under
`sdnq_b200/` imports this file: it exists so that
  * `tests/golden/generate*.py` can produce the committed fixtures by running the reference,
  * `bench.py` can time the reference itself (CPU-eager arm; its Triton / CUDA-eager paths on the same B200),
  * `tests/test_level_b_binding_gpu.py` can run the reference's own SDNQLinear.forward over this library's C ABI.

Recipe: the reference imports diffusers / accelerate at module top (quantizer.py:7-12) and neither is
installed in this image, so minimal stand-in modules are put into sys.modules first.  Flags that the reference resolves at import
time (SDNQ_DEVICE, SDNQ_USE_TORCH_COMPILE, SDNQ_USE_TRITON_MM, ...) must be in os.environ before the first call.
"""
import contextlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref"), "/root/reference/src")


def reference_root():
    for root in CANDIDATES:
        if os.path.isfile(os.path.join(root, "sdnq", "__init__.py")):
            return root
    return None


def load_reference(**env):
    """-> the reference's top-level module.  `env`: flag values to put into os.environ before the import (first call only)."""
    if "sdnq" in sys.modules:
        return sys.modules["sdnq"]
    root = reference_root()
    if root is None:
        raise RuntimeError("reference sources not present: neither oracle/_ref/sdnq (python oracle/build_ref.py) nor /root/reference/src")
    for k, v in env.items():
        os.environ[k] = str(v)
    os.environ.setdefault("SDNQ_USE_TORCH_COMPILE", "0")
    import transformers.quantizers  # noqa: F401  (must precede the accelerate stand-in: it probes accelerate with find_spec)

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

    if "diffusers" not in sys.modules:
        mod("diffusers", __version__="0.40.0", __path__=[])
        mod("diffusers.quantizers", __path__=[])
        mod("diffusers.quantizers.base", DiffusersQuantizer=DiffusersQuantizer)
        mod("diffusers.quantizers.quantization_config", QuantizationConfigMixin=QuantizationConfigMixin)
        mod("diffusers.quantizers.auto", AUTO_QUANTIZER_MAPPING={}, AUTO_QUANTIZATION_CONFIG_MAPPING={})
        mod("diffusers.utils", get_module_from_name=get_module_from_name)
    if "accelerate" not in sys.modules:
        mod("accelerate", init_empty_weights=contextlib.nullcontext)
    sys.path.insert(0, root)
    import sdnq
    return sdnq
