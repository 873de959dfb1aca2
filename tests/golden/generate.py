"""Generate the committed golden fixtures from the upstream reference.

Run in the authoring container only (needs /root/reference):

    SDNQ_USE_CONTIGUOUS_MM=0 SDNQ_ALLOW_FP8_MM=1 SDNQ_USE_TORCH_COMPILE=0 \
        python tests/golden/generate.py

The two env flags put the CPU-eager reference into the flag state it resolves
to on a B200: K-major B operand / transposed SVD
factors, fp8 matmul allowed (torch._scaled_mm runs on CPU in torch 2.11).
Everything written here is *data produced by running the unmodified reference*;
no reference source is copied.  bf16 tensors are stored as uint16 bit
patterns, fp8 tensors as uint8 bit patterns.

Outputs (all under tests/golden/):
    dtype_table.json      the reference's storage-dtype table (numeric attrs)
    pack_kat.npz          pack_int / unpack_int known answers, every width 1..15
    float_tables.npz      unpack_float decode of every code for <=8-bit minifloats
                          + pack_float encodings of a fixed value sweep
    layer_<name>.npz      per-config quantised layer + input + every
                          intermediate of the forward + output
    hadamard.npz          get_hadamard matrices + rotate_hadamard outputs
"""
import copy
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_loader import load_reference  # noqa: E402

assert os.environ.get("SDNQ_USE_CONTIGUOUS_MM") == "0" and os.environ.get("SDNQ_ALLOW_FP8_MM") == "1", __doc__
sdnq = load_reference()
from sdnq import SDNQConfig  # noqa: E402
from sdnq.common import dtype_dict  # noqa: E402
from sdnq.packed_float import pack_float, unpack_float  # noqa: E402
from sdnq.packed_int import pack_int, unpack_int  # noqa: E402
from sdnq.quant_utils import get_hadamard, rotate_hadamard  # noqa: E402
from sdnq.quantizer import sdnq_quantize_layer  # noqa: E402
from sdnq.layers.linear import linear_int8, linear_uint8, linear_fp8  # noqa: E402


def to_np(t):
    """torch tensor -> numpy, preserving bits for bf16 / fp8 / bool."""
    if t is None:
        return None
    t = t.detach().cpu()
    if t.dtype == torch.bfloat16:
        return t.contiguous().view(torch.int16).numpy().view(np.uint16)
    if t.dtype in (torch.float8_e4m3fn, torch.float8_e5m2):
        return t.contiguous().view(torch.uint8).numpy()
    if t.dtype == torch.uint16:
        return t.contiguous().view(torch.int16).numpy().view(np.uint16)
    if t.dtype == torch.uint32:
        return t.contiguous().view(torch.int32).numpy().view(np.uint32)
    return t.contiguous().numpy()


def tinfo(t):
    if t is None:
        return None
    return {"dtype": str(t.dtype).replace("torch.", ""), "shape": list(t.shape), "stride": list(t.stride())}


# --------------------------------------------------------------------------- dtype table
def dump_dtype_table():
    table = {}
    ident = {}
    for name, entry in dtype_dict.items():
        if id(entry) in ident:
            table[name] = {"alias_of": ident[id(entry)]}
            continue
        ident[id(entry)] = name
        row = {}
        for k, v in entry.items():
            row[k] = str(v).replace("torch.", "") if isinstance(v, torch.dtype) else v
        table[name] = row
    with open(os.path.join(HERE, "dtype_table.json"), "w") as f:
        json.dump(table, f, indent=0, sort_keys=True)


# --------------------------------------------------------------------------- pack KATs
def dump_pack_kat():
    g = torch.Generator().manual_seed(1234)
    out = {}
    for bits in list(range(1, 8)) + list(range(9, 16)):
        name = f"uint{bits}"
        n = 16 * 24
        if bits == 1:
            codes = torch.randint(0, 2, (n,), generator=g).to(torch.bool)
        else:
            codes = torch.randint(0, 2 ** bits, (n,), generator=g).to(dtype_dict[name]["storage_dtype"])
        packed = pack_int(codes, name)
        rt = unpack_int(packed, name, codes.shape)
        assert torch.equal(rt.to(torch.int32), codes.to(torch.int32))
        out[f"{name}_codes"] = codes.to(torch.int32).numpy()
        out[f"{name}_packed"] = to_np(packed) if packed.dtype != torch.int64 else packed.numpy()
        out[f"{name}_packed_shape"] = np.array(packed.shape)
        if bits > 1:
            sname = f"int{bits}"
            scodes = (codes.to(torch.int32) + dtype_dict[sname]["min"]).to(dtype_dict[sname]["torch_dtype"])
            spacked = pack_int(scodes, sname)
            assert torch.equal(spacked, packed)
            srt = unpack_int(spacked, sname, scodes.shape)
            assert torch.equal(srt, scodes)
    # the hand-checkable vectors of SURVEY.md 8c
    kat = {
        "uint4": [0, 1], "uint2": [0, 1, 2, 3], "uint3": list(range(8)),
        "uint5": [11, 16, 21, 26, 31, 4, 9, 14], "uint6": [11, 48, 21, 58],
        "uint7": [11, 48, 85, 122, 31, 68, 105, 14],
    }
    for name, v in kat.items():
        out[f"kat_{name}_in"] = np.array(v, dtype=np.int32)
        out[f"kat_{name}_out"] = pack_int(torch.tensor(v, dtype=torch.uint8), name).numpy().reshape(-1)
    np.savez_compressed(os.path.join(HERE, "pack_kat.npz"), **out)


# --------------------------------------------------------------------------- float tables
def dump_float_tables():
    out = {}
    sweep = torch.cat([
        torch.linspace(-520, 520, 2081),
        torch.tensor([0.0, -0.0, 1e-8, -1e-8, 2.0 ** -9, 2.0 ** -6, 0.3, 0.75, 1.5, 2.5, 3.5, 5.0, 6.0, 7.0, 448.0, 480.0]),
        torch.randn(512, generator=torch.Generator().manual_seed(7)) * 3,
    ]).to(torch.float32)
    out["sweep"] = sweep.numpy()
    names = []
    for name, e in dtype_dict.items():
        if "alias" in name or e["is_integer"] or not e["is_packed"] or e["num_bits"] > 8:
            continue
        if not name.startswith("float"):
            continue
        names.append(name)
        bits = e["num_bits"]
        codes = torch.arange(2 ** bits, dtype=torch.int32)
        n = codes.numel()
        pad = (-n) % 8
        codes_p = torch.cat([codes, torch.zeros(pad, dtype=torch.int32)])
        if bits == 8:
            packed = codes_p.to(torch.uint8)
        else:
            uname = f"uint{bits}"
            packed = pack_int(codes_p.to(torch.bool if bits == 1 else torch.uint8), uname)
        vals = unpack_float(packed, name, codes_p.shape)[:n]
        out[f"{name}_decode"] = vals.numpy()
        # encode a clamped sweep (the quantiser clamps before pack_float)
        clamped = sweep.clamp(e["min"], e["max"])
        enc = pack_float(clamped[: (clamped.numel() // 8) * 8], name)
        dec = unpack_float(enc, name, (clamped.numel() // 8 * 8,))
        out[f"{name}_encode_roundtrip"] = dec.numpy()
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "float_tables.npz"), **out)


# --------------------------------------------------------------------------- hadamard
def dump_hadamard():
    out = {}
    g = torch.Generator().manual_seed(99)
    for n in (4, 8, 16, 32, 64, 128, 256, 512):
        H32 = get_hadamard(n, dtype=torch.float32, device=torch.device("cpu"))
        Hbf = get_hadamard(n, dtype=torch.bfloat16, device=torch.device("cpu"))
        out[f"H{n}_f32"] = H32.contiguous().numpy()
        out[f"H{n}_bf16"] = to_np(Hbf)
        x = torch.randn(6, 2 * n, generator=g)
        out[f"x{n}_f32"] = x.numpy()
        out[f"y{n}_f32"] = rotate_hadamard(x, hadamard=H32).numpy()
        xb = x.to(torch.bfloat16)
        out[f"y{n}_bf16"] = to_np(rotate_hadamard(xb, hadamard=Hbf))
    np.savez_compressed(os.path.join(HERE, "hadamard.npz"), **out)


# --------------------------------------------------------------------------- layers
LAYER_CASES = {
    # name: (K, N, M, bias, SDNQConfig kwargs)
    "c1_int8_rowwise_dequant":      (256, 128, 48, True,  dict(weights_dtype="int8", group_size=-1)),
    "c2_int8_w8a8":                 (256, 128, 48, True,  dict(weights_dtype="int8", use_quantized_matmul=True)),
    "c2_int8_w8a8_nobias":          (256, 128, 40, False, dict(weights_dtype="int8", use_quantized_matmul=True)),
    "uint8_w8a8":                   (256, 128, 48, True,  dict(weights_dtype="uint8", use_quantized_matmul=True)),
    "c3_fp8_hadamard_w8a8":         (512, 128, 48, True,  dict(weights_dtype="float8_e4m3fn", use_quantized_matmul=True, use_hadamard=True, hadamard_group_size=256)),
    "c3b_fp8_g256_hadamard_w8a8":   (512, 128, 48, True,  dict(weights_dtype="float8_e4m3fn", group_size=256, use_quantized_matmul=True, use_hadamard=True, hadamard_group_size=256)),
    "fp8_dequant":                  (256, 128, 48, True,  dict(weights_dtype="float8_e4m3fn")),
    "c4_int4_g128_svd_dequant":     (256, 128, 48, True,  dict(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32)),
    "int4_g128_svd_w8a8":           (256, 128, 48, True,  dict(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32, use_quantized_matmul=True)),
    "int4_rowwise_w8a8":            (256, 128, 48, True,  dict(weights_dtype="int4", group_size=-1, use_quantized_matmul=True)),
    "uint4_auto_dequant":           (256, 128, 48, True,  dict(weights_dtype="uint4")),
    "uint4_auto_w8a8":              (256, 128, 48, True,  dict(weights_dtype="uint4", use_quantized_matmul=True)),
    "uint4_rowwise_w8a8":           (256, 128, 48, True,  dict(weights_dtype="uint4", group_size=-1, use_quantized_matmul=True)),
    "float6_e3m2_w8a8":             (256, 128, 48, True,  dict(weights_dtype="float6_e3m2fn", use_quantized_matmul=True)),
    "int8_hadamard128_w8a8":        (384, 128, 48, True,  dict(weights_dtype="int8", use_quantized_matmul=True, use_hadamard=True)),
    "int8_svd_w8a8":                (256, 128, 48, True,  dict(weights_dtype="int8", use_quantized_matmul=True, use_svd=True, svd_rank=16)),
    "int8_hadamard_svd_dequant":    (256, 128, 48, True,  dict(weights_dtype="int8", use_hadamard=True, use_svd=True, svd_rank=16)),
    "uint3_dequant":                (256, 128, 40, False, dict(weights_dtype="uint3")),
    "int5_dequant":                 (256, 128, 40, True,  dict(weights_dtype="int5", group_size=32)),
    "int6_rowwise_w8a8":            (256, 128, 40, True,  dict(weights_dtype="int6", use_quantized_matmul=True)),
    "uint7_dequant":                (256, 128, 40, True,  dict(weights_dtype="uint7", group_size=64)),
    "int7_dequant":                 (256, 128, 40, True,  dict(weights_dtype="int7", group_size=-1)),
    "int2_dequant":                 (256, 128, 40, True,  dict(weights_dtype="int2")),
    "uint2_hadamard_dequant":       (256, 128, 40, True,  dict(weights_dtype="uint2", use_hadamard=True)),
    "uint1_dequant":                (256, 128, 40, True,  dict(weights_dtype="uint1", group_size=32)),
    "float4_e2m1_dequant":          (256, 128, 40, True,  dict(weights_dtype="float4_e2m1fn", group_size=32)),
    "float8_e4m3fn_sdnq_dequant":   (256, 128, 40, True,  dict(weights_dtype="float8_e4m3fn_sdnq", group_size=-1)),
    "float5_e2m2_w8a8":             (256, 128, 40, True,  dict(weights_dtype="float5_e2m2fn", use_quantized_matmul=True)),
    "float7_e3m3_rowwise_w8a8":     (256, 128, 40, True,  dict(weights_dtype="float7_e3m3fn", group_size=-1, use_quantized_matmul=True)),
    "float4_e2m2fnu_dequant":       (256, 128, 40, True,  dict(weights_dtype="float4_e2m2fnu", group_size=64)),
    "uint4_codebook_dequant":       (256, 128, 40, True,  dict(weights_dtype="uint4", use_codebook=True, group_size=64)),
    "int8_w8a8_small_m":            (256, 128, 8,  True,  dict(weights_dtype="int8", use_quantized_matmul=True)),
    "int8_tensorwise_dequant":      (256, 128, 40, True,  dict(weights_dtype="int8", group_size=-2)),
    "int8_w8a8_outliers":           (512, 256, 64, True,  dict(weights_dtype="int8", use_quantized_matmul=True)),
    # round 2: float8_e5m2 weights under the e4m3 matmul (mixed e4m3 x e5m2 operands), its small-M branch and its dequant path
    "fp8_e5m2_w8a8":                (256, 128, 48, True,  dict(weights_dtype="float8_e5m2", use_quantized_matmul=True)),
    "fp8_e5m2_w8a8_small_m":        (256, 128, 8,  True,  dict(weights_dtype="float8_e5m2", use_quantized_matmul=True)),
    "fp8_e5m2_dequant":             (256, 128, 40, True,  dict(weights_dtype="float8_e5m2")),
    # round 2: rows < 32 of packed / group-wise layers (the K5p kernel) against the reference's dequantise + F.linear
    "c4_int4_g128_svd_dequant_small_m": (256, 128, 8, True, dict(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32)),
    "uint4_auto_w8a8_small_m":      (256, 128, 8,  True,  dict(weights_dtype="uint4", use_quantized_matmul=True)),
    "float6_e3m2_w8a8_small_m":     (256, 128, 5,  True,  dict(weights_dtype="float6_e3m2fn", use_quantized_matmul=True)),
    "int5_dequant_small_m":         (256, 128, 4,  True,  dict(weights_dtype="int5", group_size=32)),
    "uint2_hadamard_dequant_small_m": (256, 128, 31, True, dict(weights_dtype="uint2", use_hadamard=True)),
    "int4_g128_svd_w8a8_small_m":   (256, 128, 16, True,  dict(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32, use_quantized_matmul=True)),
}


def run_layer_case(name, K, N, M, bias, cfg):
    torch.manual_seed(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    lin = torch.nn.Linear(K, N, bias=bias).to(torch.bfloat16)
    x = torch.randn(M, K).to(torch.bfloat16)
    if "outliers" in name:
        x[:, ::97] *= 20
        x[3] = 0.0  # an all-zero row: the reference produces NaN there (0/0), unguarded
    w_orig = lin.weight.detach().clone()
    b_orig = None if lin.bias is None else lin.bias.detach().clone()
    layer = sdnq_quantize_layer(copy.deepcopy(lin), SDNQConfig(**cfg))[0]
    d = layer.sdnq_dequantizer
    arrays = {"x": to_np(x), "w_orig": to_np(w_orig)}
    if b_orig is not None:
        arrays["bias"] = to_np(b_orig)
    meta = {"name": name, "K": K, "N": N, "M": M, "config": cfg, "tensors": {}}
    for key in ("weight", "scale", "zero_point", "svd_up", "svd_down"):
        t = getattr(layer, key)
        meta["tensors"][key] = tinfo(t)
        if t is not None:
            # store the *physical* bytes of a K-major weight (stride (1,K)) as its [N,K] transpose
            tt = t.detach()
            if tt.ndim == 2 and not tt.is_contiguous() and tt.t().is_contiguous():
                arrays[key + "__T"] = to_np(tt.t())
            else:
                arrays[key] = to_np(tt)
    meta["dequantizer"] = {
        k: (str(v).replace("torch.", "") if isinstance(v, torch.dtype) else (list(v) if isinstance(v, (torch.Size, tuple)) else v))
        for k, v in d.__dict__.items()
    }
    meta["forward_func"] = layer.forward_func.__name__
    with torch.no_grad():
        y = layer(x)
        arrays["y"] = to_np(y)
        W = d(layer.weight, layer.scale, layer.zero_point, layer.svd_up, layer.svd_down, skip_quantized_matmul=d.use_quantized_matmul)
        arrays["w_dequant"] = to_np(W)
        meta["w_dequant"] = tinfo(W)
        if d.use_quantized_matmul and M >= 32:
            hadamard = get_hadamard(d.hadamard_group_size, dtype=x.dtype, device=x.device) if d.use_hadamard else None
            zp = layer.zero_point
            if d.re_quantize_for_matmul:
                rq = d.re_quantize_matmul(layer.weight, layer.scale, zero_point=layer.zero_point)
                arrays["rq_weight__T"] = to_np(rq[0].t())
                arrays["rq_scale"] = to_np(rq[1])
                if len(rq) == 3:
                    arrays["rq_zero_point"] = to_np(rq[2])
                    zp = rq[2]
                else:
                    zp = None
                wq, sw, qshape = rq[0], rq[1], None
            else:
                wq, sw = layer.weight, layer.scale
                qshape = d.quantized_weight_shape if d.is_packed else None
            fn = layer.forward_func.__name__
            if fn.endswith("_int8_matmul"):
                r = linear_int8.get_int8_matmul_inputs(x, wq, sw, bias=layer.bias, svd_up=layer.svd_up, svd_down=layer.svd_down,
                                                       zero_point=zp, hadamard=hadamard, quantized_weight_shape=qshape, weights_dtype=d.weights_dtype)
            elif fn.endswith("_uint8_matmul"):
                r = linear_uint8.get_uint8_matmul_inputs(x, wq, sw, zp, bias=layer.bias, svd_up=layer.svd_up, svd_down=layer.svd_down,
                                                         hadamard=hadamard, quantized_weight_shape=qshape, weights_dtype=d.weights_dtype)
            elif fn.endswith("_fp8_matmul"):
                r = linear_fp8.get_fp8_matmul_inputs(x, wq, sw, bias=layer.bias, svd_up=layer.svd_up, svd_down=layer.svd_down,
                                                     hadamard=hadamard, quantized_weight_shape=qshape, weights_dtype=d.weights_dtype)
            else:
                r = None
            if r is not None:
                xq, wmm, sx, swmm, mm_bias = r[:5]
                arrays["mm_xq"] = to_np(xq)
                arrays["mm_sx"] = to_np(sx)
                arrays["mm_wq__T"] = to_np(wmm.t())
                arrays["mm_sw"] = to_np(swmm)
                if mm_bias is not None:
                    arrays["mm_bias"] = to_np(mm_bias)
                    meta["mm_bias"] = tinfo(mm_bias)
                if hadamard is not None:
                    arrays["x_rot"] = to_np(rotate_hadamard(x, hadamard=hadamard))
    arrays["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, f"layer_{name}.npz"), **arrays)
    nan = int(np.isnan(y.float().numpy()).sum())
    print(f"{name:34s} fwd={layer.forward_func.__name__:40s} w={tuple(layer.weight.shape)} gs={d.group_size} requant={d.re_quantize_for_matmul} nan={nan}")


def main():
    only = [n for n in os.environ.get("SDNQ_GOLDEN_ONLY", "").split(",") if n]      # regenerate just the named layer cases
    if not only:
        dump_dtype_table()
        dump_pack_kat()
        dump_float_tables()
        dump_hadamard()
        dump_policy_tables()
    for name, (K, N, M, bias, cfg) in LAYER_CASES.items():
        if only and name not in only:
            continue
        run_layer_case(name, K, N, M, bias, cfg)
    total = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith((".npz", ".json")))
    print("fixtures bytes:", total)


# --------------------------------------------------------------------------- policy tables (data, not code)
def dump_policy_tables():
    """weights_dtype_order and the per-architecture skip-key tables (reference common.py:302-334, 383-526).
    Pure configuration data; sdnq_b200/skip_keys.json is this dump and tests/test_host_api.py checks they agree."""
    from sdnq.common import common_skip_keys, module_skip_keys_dict, weights_dtype_order
    seen = {}
    models, aliases = {}, {}
    for name, entry in module_skip_keys_dict.items():
        if id(entry) in seen:
            aliases[name] = seen[id(entry)]
            continue
        seen[id(entry)] = name
        models[name] = {"modules_to_not_convert": list(entry[0]), "modules_dtype_dict": entry[1], "modules_to_not_use_matmul": entry[2]}
    tables = {"weights_dtype_order": list(weights_dtype_order), "common_skip_keys": list(common_skip_keys), "models": models, "aliases": aliases}
    with open(os.path.join(HERE, "policy_tables.json"), "w") as f:
        json.dump(tables, f, indent=1, sort_keys=True)
    pkg = os.path.join(os.path.dirname(os.path.dirname(HERE)), "sdnq_b200", "skip_keys.json")
    with open(pkg, "w") as f:
        json.dump({k: tables[k] for k in ("common_skip_keys", "models", "aliases")}, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    if os.environ.get("SDNQ_GOLDEN_POLICY_ONLY"):
        dump_policy_tables()
    else:
        main()
