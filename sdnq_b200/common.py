"""Storage-dtype table and layer-type sets of the SDNQ surface.

The reference spells this table out literally (reference common.py:16-334); here it is generated from the
naming rule and pinned entry-by-entry against a dump of the reference's table
(tests/golden/dtype_table.json, tests/test_host_api.py).  Known upstream quirk kept on purpose: uint9..uint15
list `max = 2**bits` (not 2**bits - 1), and results must match the reference's.
"""
import torch

sdnq_version = "0.2.5"          # checkpoint-format version this implementation is compatible with
sdnq_keys = {"weight", "scale", "zero_point", "svd_up", "svd_down"}     # reference common.py:8


def _int_entry(bits: int, unsigned: bool) -> dict:
    name = f"{'u' if unsigned else ''}int{bits}"
    native = {8: (torch.uint8, torch.int8), 16: (torch.uint16, torch.int16), 32: (torch.uint32, torch.int32)}
    if bits in native:
        td = native[bits][0 if unsigned else 1]
        target, torch_dtype, storage, packed = td, td, td, False
    elif bits == 1:
        target, torch_dtype, storage, packed = torch.bool, torch.bool, torch.bool, True
    elif bits < 8:
        target, torch_dtype, storage, packed = name, (torch.uint8 if unsigned else torch.int8), torch.uint8, True
    else:
        target, torch_dtype, storage, packed = name, torch.int16, torch.int16, True
    if unsigned:
        lo, hi = 0, (2 ** bits if 9 <= bits <= 15 else 2 ** bits - 1)
    else:
        lo, hi = -(2 ** (bits - 1)), 2 ** (bits - 1) - 1
    return {"min": lo, "max": hi, "num_bits": bits, "sign": 0 if unsigned else 1, "exponent": 0,
            "mantissa": bits if unsigned else bits - 1, "target_dtype": target, "torch_dtype": torch_dtype,
            "storage_dtype": storage, "is_unsigned": unsigned, "is_integer": True, "is_packed": packed}


def _minifloat_entry(bits: int, exponent: int, mantissa: int, unsigned: bool) -> dict:
    bias = 2 ** (exponent - 1) - 1
    mx = (2.0 - 2.0 ** (-mantissa)) * 2.0 ** (2 ** exponent - 1 - bias)
    storage = torch.uint8 if bits <= 8 else (torch.uint16 if bits == 16 else torch.int16)
    return {"min": 0 if unsigned else -mx, "max": mx, "num_bits": bits, "sign": 0 if unsigned else 1,
            "exponent": exponent, "mantissa": mantissa, "target_dtype": f"fp{bits}", "torch_dtype": torch.float32,
            "storage_dtype": storage, "is_unsigned": unsigned, "is_integer": False, "is_packed": True}


def _native_float(td, mx, exponent, mantissa, target=None) -> dict:
    return {"min": -mx, "max": mx, "num_bits": 1 + exponent + mantissa, "sign": 1, "exponent": exponent, "mantissa": mantissa,
            "target_dtype": td if target is None else target, "torch_dtype": td, "storage_dtype": td, "is_unsigned": False,
            "is_integer": False, "is_packed": False}


def _build_dtype_dict() -> dict:
    d = {}
    for bits in (32, 16, 8, 15, 14, 13, 12, 11, 10, 9, 7, 6, 5, 4, 3, 2):
        d[f"int{bits}"] = _int_entry(bits, False)
    for bits in (32, 16, 8, 15, 14, 13, 12, 11, 10, 9, 7, 6, 5, 4, 3, 2, 1):
        d[f"uint{bits}"] = _int_entry(bits, True)
    d["float32"] = _native_float(torch.float32, 3.40282e+38, 8, 23)
    d["bfloat16"] = _native_float(torch.bfloat16, 3.38953e+38, 8, 7)
    d["float16"] = _native_float(torch.float16, 65504.0, 5, 10)
    d["float8_e4m3fn"] = _native_float(torch.float8_e4m3fn, 448.0, 4, 3)
    d["float8_e5m2"] = _native_float(torch.float8_e5m2, 57344.0, 5, 2)
    for unsigned in (False, True):
        for bits in range(16, 0, -1):
            for exponent in range(1, 6):
                mantissa = bits - exponent - (0 if unsigned else 1)
                if mantissa < 0 or (not unsigned and bits < 2):
                    continue
                name = f"float{bits}_e{exponent}m{mantissa}{'fnu' if unsigned else 'fn'}"
                if name == "float8_e4m3fn":
                    name = "float8_e4m3fn_sdnq"      # the packed "every code finite" variant (max 480)
                d[name] = _minifloat_entry(bits, exponent, mantissa, unsigned)
    aliases = {"fp32": "float32", "bf16": "bfloat16", "fp16": "float16", "fp8": "float8_e4m3fn", "int1": "uint1", "bool": "uint1"}
    signed_default = {15: (5, 9), 14: (5, 8), 13: (5, 7), 12: (5, 6), 11: (5, 5), 10: (5, 4), 9: (4, 4), 7: (3, 3), 6: (3, 2),
                      5: (2, 2), 4: (2, 1), 3: (1, 1), 2: (1, 0)}
    for bits, (e, m) in signed_default.items():
        aliases[f"fp{bits}"] = f"float{bits}_e{e}m{m}fn"
    unsigned_default = {16: (5, 11), 15: (5, 10), 14: (5, 9), 13: (5, 8), 12: (5, 7), 11: (5, 6), 10: (5, 5), 9: (4, 5), 8: (4, 4),
                        7: (3, 4), 6: (3, 3), 5: (2, 3), 4: (2, 2), 3: (1, 2), 2: (1, 1), 1: (1, 0)}
    for bits, (e, m) in unsigned_default.items():
        aliases[f"ufp{bits}"] = f"float{bits}_e{e}m{m}fnu"
    aliases["fp1"] = "float1_e1m0fnu"
    for alias, target in aliases.items():
        d[alias] = d[target]
    if hasattr(torch, "float8_e8m0fnu"):
        d["float8_e8m0fnu"] = dict(_native_float(torch.float8_e8m0fnu, 1.70141e+38, 8, 0, target="fp8"), num_bits=8)
    if hasattr(torch, "float8_e4m3fnuz"):
        d["float8_e4m3fnuz"] = _native_float(torch.float8_e4m3fnuz, 240.0, 4, 3, target="fp8")
    if hasattr(torch, "float8_e5m2fnuz"):
        d["float8_e5m2fnuz"] = _native_float(torch.float8_e5m2fnuz, 57344.0, 5, 2, target="fp8")
    return d


dtype_dict = _build_dtype_dict()

linear_types = {"Linear", "SDNQLinear"}
embedding_types = {"Embedding", "SDNQEmbedding", "Gemma4TextScaledWordEmbedding"}
conv_types = {"Conv1d", "Conv2d", "Conv3d", "SDNQConv1d", "SDNQConv2d", "SDNQConv3d"}
conv_transpose_types = {"ConvTranspose1d", "ConvTranspose2d", "ConvTranspose3d", "SDNQConvTranspose1d", "SDNQConvTranspose2d", "SDNQConvTranspose3d"}
allowed_types = set.union(linear_types, embedding_types, conv_types, conv_transpose_types)

accepted_weight_dtypes = set(dtype_dict.keys())
accepted_matmul_dtypes = {"int8", "uint8", "fp8", "fp16", "float8_e4m3fn", "float16"}


def _build_weights_dtype_order() -> list:
    """Ascending "precision" order used by dynamic quantisation to pick the next dtype (reference common.py:302-334):
    per bit-width: signed int, signed minifloats e1..e5, then unsigned int, unsigned minifloats e1..e5.  At 8 and 16
    bits the native torch types come right after the integer."""
    order = []
    for bits in range(1, 17):
        for unsigned in (False, True):
            if bits == 1 and not unsigned:
                continue
            order.append(f"{'u' if unsigned else ''}int{bits}")
            if not unsigned and bits == 8:
                order += ["float8_e4m3fn", "float8_e5m2"]
            if not unsigned and bits == 16:
                order.append("float16")
            for exponent in range(1, 6):
                mantissa = bits - exponent - (0 if unsigned else 1)
                if mantissa < 0:
                    continue
                name = f"float{bits}_e{exponent}m{mantissa}{'fnu' if unsigned else 'fn'}"
                order.append("float8_e4m3fn_sdnq" if name == "float8_e4m3fn" else name)
    return order


weights_dtype_order = _build_weights_dtype_order()


def _load_skip_keys():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "skip_keys.json")) as f:
        raw = json.load(f)
    table = {name: [e["modules_to_not_convert"], e["modules_dtype_dict"], e["modules_to_not_use_matmul"]] for name, e in raw["models"].items()}
    for alias, target in raw["aliases"].items():
        table[alias] = table[target]
    return tuple(raw["common_skip_keys"]), table


# per-architecture lists of modules that stay unquantised / keep bf16 matmul (policy data, see skip_keys.json)
common_skip_keys, module_skip_keys_dict = _load_skip_keys()
