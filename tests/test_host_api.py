"""CPU tests of the host-side surface (config, policy helpers, packers, quantise-time layout rules) against
reference-generated fixtures.  No kernels are launched here."""
import copy
import json
import os

import numpy as np
import pytest
import torch

from tests.util import GOLDEN, LAYER_FILES, LAYER_IDS, fixture_tensors


def test_dtype_table_equals_reference_dump():
    from sdnq_b200.common import dtype_dict
    table = json.load(open(os.path.join(GOLDEN, "dtype_table.json")))
    assert set(table) == set(dtype_dict)
    for name, row in table.items():
        if "alias_of" in row:
            assert dtype_dict[name] is dtype_dict[row["alias_of"]], name
            continue
        for key, ref in row.items():
            mine = dtype_dict[name][key]
            mine = str(mine).replace("torch.", "") if isinstance(mine, torch.dtype) else mine
            assert mine == ref or (isinstance(ref, float) and abs(mine - ref) <= 1e-5 * abs(ref)), (name, key, mine, ref)


def test_policy_tables_equal_reference_dump():
    from sdnq_b200.common import common_skip_keys, module_skip_keys_dict, weights_dtype_order
    ref = json.load(open(os.path.join(GOLDEN, "policy_tables.json")))
    assert weights_dtype_order == ref["weights_dtype_order"]
    assert list(common_skip_keys) == ref["common_skip_keys"]
    for name, entry in ref["models"].items():
        assert module_skip_keys_dict[name] == [entry["modules_to_not_convert"], entry["modules_dtype_dict"], entry["modules_to_not_use_matmul"]]
    for alias, target in ref["aliases"].items():
        assert module_skip_keys_dict[alias] is module_skip_keys_dict[target]


def test_check_param_name_in_rules():
    from sdnq_b200.utils import check_param_name_in
    keys = [".proj_out", "norm_out", "time_text_embed", "transformer_blocks.0.norm*"]
    assert check_param_name_in("proj_out.weight", keys) == ".proj_out"                       # leading dot = top-level prefix only
    assert check_param_name_in("single_transformer_blocks.3.proj_out.weight", keys) is None
    assert check_param_name_in("blocks.1.norm_out.linear.weight", keys) == "norm_out"        # whole dotted segment
    assert check_param_name_in("blocks.1.norm_out2.weight", keys) is None
    assert check_param_name_in("transformer_blocks.0.norm1.linear.weight", keys) == "transformer_blocks.0.norm*"
    assert check_param_name_in("transformer_blocks.10.norm1.linear.weight", keys) is None


def test_config_validation_and_json_roundtrip():
    from sdnq_b200 import SDNQConfig
    cfg = SDNQConfig(weights_dtype="uint4", use_quantized_matmul=True, modules_to_not_convert="lm_head", modules_dtype_dict={"int8": "proj"})
    d = cfg.to_dict()
    assert d["modules_to_not_convert"] == ["lm_head"] and d["modules_dtype_dict"] == {"int8": ["proj"]} and d["quant_method"] == "sdnq"
    again = SDNQConfig(**{k: v for k, v in d.items() if k not in ("is_integer", "is_unsigned", "quant_method")})
    assert again.to_dict() == d
    with pytest.raises(ValueError):
        SDNQConfig(weights_dtype="int17")
    with pytest.raises(ValueError):
        SDNQConfig(weights_dtype="int4", use_codebook=True)
    with pytest.raises(ValueError):
        SDNQConfig(quantized_matmul_dtype="int4")


@pytest.mark.parametrize("bits", list(range(1, 8)) + list(range(9, 16)))
def test_pack_int_matches_reference_vectors(bits):
    from sdnq_b200.packing import pack_int, unpack_int
    z = np.load(os.path.join(GOLDEN, "pack_kat.npz"))
    codes, packed = z[f"uint{bits}_codes"], z[f"uint{bits}_packed"]
    mine = pack_int(torch.from_numpy(codes), f"uint{bits}")
    assert list(mine.shape) == list(z[f"uint{bits}_packed_shape"])
    assert np.array_equal(mine.numpy().astype(np.int64) & 0xFFFF, packed.astype(np.int64) & 0xFFFF)
    assert np.array_equal(unpack_int(mine, f"uint{bits}", codes.shape).numpy().astype(np.int64), codes)
    if bits > 1:
        signed = torch.from_numpy(codes) - 2 ** (bits - 1)
        assert torch.equal(pack_int(signed, f"int{bits}"), mine)
        assert torch.equal(unpack_int(mine, f"int{bits}", codes.shape).to(torch.int64), signed.to(torch.int64))


def test_pack_float_roundtrip_matches_reference():
    from sdnq_b200.common import dtype_dict
    from sdnq_b200.packing import pack_float, unpack_float
    z = np.load(os.path.join(GOLDEN, "float_tables.npz"))
    sweep = torch.from_numpy(z["sweep"])
    n = sweep.numel() // 8 * 8
    for name in [str(s) for s in z["names"]]:
        info = dtype_dict[name]
        clamped = sweep.clamp(info["min"], info["max"])[:n]
        mine = unpack_float(pack_float(clamped, name), name, (n,))
        ref = z[f"{name}_encode_roundtrip"]
        assert np.array_equal(mine.numpy().view(np.uint32), ref.view(np.uint32)), name


@pytest.mark.parametrize("path", LAYER_FILES, ids=LAYER_IDS)
def test_quantize_layer_reproduces_reference_tensors(path):
    """same float weight + same config -> same stored tensors and metadata as the reference produced."""
    from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
    t, z, meta = fixture_tensors(path)
    cfg = meta["config"]
    lin = torch.nn.Linear(meta["K"], meta["N"], bias=t["bias"] is not None).to(torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(t["w_orig"])
        if t["bias"] is not None:
            lin.bias.copy_(t["bias"])
    torch.manual_seed(0)
    layer, _ = sdnq_quantize_layer(copy.deepcopy(lin), SDNQConfig(**cfg))
    d, ref = layer.sdnq_dequantizer, meta["dequantizer"]
    assert layer.forward_func.__name__ == meta["forward_func"]
    for key in ("weights_dtype", "quantized_matmul_dtype", "group_size", "hadamard_group_size", "use_quantized_matmul", "re_quantize_for_matmul",
                "use_hadamard", "use_codebook", "is_packed", "is_unsigned", "is_integer", "is_integer_matmul", "svd_rank", "layer_class_name"):
        assert getattr(d, key) == ref[key], key
    assert list(d.quantized_weight_shape) == ref["quantized_weight_shape"] and list(d.original_shape) == ref["original_shape"]
    assert (None if d.result_shape is None else list(d.result_shape)) == ref["result_shape"]
    exact = not cfg.get("use_svd", False)       # svd_lowrank is randomised: the residual (and so the codes) differ run to run
    for key in ("weight", "scale", "zero_point", "svd_up", "svd_down"):
        mine, theirs, info = getattr(layer, key), t[key], meta["tensors"][key]
        assert (mine is None) == (theirs is None), key
        if mine is None:
            continue
        assert list(mine.shape) == info["shape"] and str(mine.dtype).replace("torch.", "") == info["dtype"], (key, mine.shape, mine.dtype, info)
        assert all(a == b or n == 1 for a, b, n in zip(mine.stride(), info["stride"], info["shape"])), (key, mine.stride(), info["stride"])
        if exact and not cfg.get("use_codebook", False):
            a = mine.detach().view(torch.uint8) if mine.dtype == torch.float8_e4m3fn else mine.detach()
            b = theirs.view(torch.uint8) if theirs.dtype == torch.float8_e4m3fn else theirs
            assert torch.equal(a.contiguous() if a.dtype != torch.bfloat16 else a.float(), b.contiguous() if b.dtype != torch.bfloat16 else b.float()), key


def test_apply_sdnq_to_module_skips_and_swaps():
    from sdnq_b200 import SDNQConfig, sdnq_post_load_quant
    from sdnq_b200.layers import SDNQLinear

    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.to_q = torch.nn.Linear(256, 256)
            self.proj_out = torch.nn.Linear(256, 256)
            self.tiny = torch.nn.Linear(16, 16)
            self.norm = torch.nn.LayerNorm(256)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.proj_out = torch.nn.Linear(256, 64)          # top-level: hit by the ".proj_out" common skip key
            self.blocks = torch.nn.ModuleList([Block(), Block()])

    net = sdnq_post_load_quant(Net(), weights_dtype="int8", use_quantized_matmul=True)
    assert not isinstance(net.proj_out, SDNQLinear)
    assert isinstance(net.blocks[0].to_q, SDNQLinear) and isinstance(net.blocks[1].proj_out, SDNQLinear)
    assert not isinstance(net.blocks[0].tiny, SDNQLinear)                     # below minimum_allowed_numel
    assert "blocks.0.tiny.weight" in net.quantization_config.modules_to_not_convert
    assert net.blocks[0].to_q.weight.dtype == torch.int8 and net.blocks[0].to_q.weight.stride() == (1, 256)
    assert set(net.blocks[0].to_q.state_dict().keys()) >= {"weight", "scale", "bias"}
    with pytest.raises(RuntimeError):
        sdnq_post_load_quant(net, weights_dtype="int8")                        # already quantised


# ----------------------------------------------------------------------------------------------- convolution layers (host side)
from tests.util import CONV_FILES, CONV_IDS, build_conv_layer  # noqa: E402


@pytest.mark.parametrize("path", CONV_FILES, ids=CONV_IDS)
def test_conv_quantisation_stores_what_the_reference_stores(path):
    """sdnq_quantize_layer on Conv / ConvTranspose modules: same dequantizer metadata, forward_func and bit-identical stored
    weight / scale / zero_point as the reference (tests/golden/generate_conv.py); the asserts live in build_conv_layer."""
    layer, t, z, meta = build_conv_layer(path)
    assert type(layer).__name__ == "SDNQ" + meta["module"]


def test_forward_dispatch_table_matches_reference():
    """get_forward_func(layer class, matmul dtype, use_quantized_matmul) returns a function of the same name as the reference's for
    every combination (tests/golden/forward_dispatch.json was written by calling the reference's get_forward_func, forward.py:6-57)."""
    from sdnq_b200.forward import get_forward_func
    table = json.load(open(os.path.join(GOLDEN, "forward_dispatch.json")))
    assert len(table) >= 80
    for key, ref_name in table.items():
        cls, mm, use = key.split("|")
        assert get_forward_func(cls, mm, bool(int(use))).__name__ == ref_name, key


# ----------------------------------------------------------------------------------------------- model-level drop-in parity
import hashlib  # noqa: E402
import sys  # noqa: E402

sys.path.insert(0, GOLDEN)
import toy_model  # noqa: E402

_MODEL_PARITY = json.load(open(os.path.join(GOLDEN, "model_parity.json")))


def _describe(model):
    mods = {}
    for name, m in model.named_modules():
        if name == "":
            continue
        e = {"class": type(m).__name__}
        d = getattr(m, "sdnq_dequantizer", None)
        if d is not None:
            e["forward_func"] = m.forward_func.__name__
            e["dequantizer"] = {k: (str(v).replace("torch.", "") if isinstance(v, torch.dtype) else (list(v) if isinstance(v, (torch.Size, tuple)) else v))
                                for k, v in d.__dict__.items()}
        mods[name] = e
    tensors = {}
    for key, t in model.state_dict().items():
        tt = t.detach()
        phys = tt.t() if (tt.ndim == 2 and not tt.is_contiguous() and tt.t().is_contiguous()) else tt.contiguous()
        raw = phys.view(torch.uint8) if phys.dtype in (torch.float8_e4m3fn, torch.float8_e5m2) else phys
        data = raw.view(torch.uint8).numpy().tobytes() if raw.dtype != torch.bfloat16 else raw.view(torch.int16).numpy().tobytes()
        tensors[key] = {"dtype": str(tt.dtype).replace("torch.", ""), "shape": list(tt.shape), "stride": list(tt.stride()),
                        "sha1": hashlib.sha1(data).hexdigest()}
    return mods, tensors


@pytest.mark.parametrize("name", sorted(_MODEL_PARITY))
def test_post_load_quant_matches_reference_on_a_model(name):
    """sdnq_post_load_quant on a UNet-shaped toy model (tests/golden/toy_model.py): which modules are swapped (skip keys, size
    thresholds, quant_conv / quant_embedding, per-module dtype overrides), the wrapper class and forward function of each, the
    dequantizer metadata, every state-dict key with its dtype / shape / stride and the bytes themselves, and the bookkeeping lists of
    the quantization config -- all equal to what the reference produced (tests/golden/generate_model.py)."""
    from sdnq_b200 import sdnq_post_load_quant
    ref = _MODEL_PARITY[name]
    model = sdnq_post_load_quant(toy_model.build(), **ref["config"])
    mods, tensors = _describe(model)
    assert sorted(mods) == sorted(ref["modules"])
    for mname, e in ref["modules"].items():
        got = mods[mname]
        assert got["class"] == e["class"], mname
        assert got.get("forward_func") == e.get("forward_func"), mname
        if "dequantizer" in e:
            for k, v in e["dequantizer"].items():
                assert got["dequantizer"].get(k) == v, (mname, k, got["dequantizer"].get(k), v)
    assert sorted(tensors) == sorted(ref["tensors"])
    for key, e in ref["tensors"].items():
        got = tensors[key]
        assert got["dtype"] == e["dtype"] and got["shape"] == e["shape"], key
        assert all(a == b or n == 1 for a, b, n in zip(got["stride"], e["stride"], e["shape"])), (key, got["stride"], e["stride"])
        assert got["sha1"] == e["sha1"], f"{key}: stored bytes differ from the reference's"
    qc = model.quantization_config
    assert sorted(qc.modules_to_not_convert) == ref["modules_to_not_convert"]
    assert {k: sorted(v) for k, v in qc.modules_dtype_dict.items()} == ref["modules_dtype_dict"]
    assert sorted(qc.modules_to_not_use_matmul) == ref["modules_to_not_use_matmul"]


_MODEL_FLIPS = json.load(open(os.path.join(GOLDEN, "model_flips.json")))


@pytest.mark.parametrize("name", sorted(_MODEL_FLIPS))
def test_apply_sdnq_options_matches_reference_on_a_model(name):
    """apply_sdnq_options_to_model (in-place matmul on / off flips of the stored layouts, loader.py:221-346) on the toy model: every
    state-dict entry (dtype, shape, stride, bytes), wrapper class, forward function and dequantizer field afterwards equals what the
    reference's own function left behind (tests/golden/generate_model.py)."""
    from sdnq_b200 import apply_sdnq_options_to_model, sdnq_post_load_quant
    ref = _MODEL_FLIPS[name]
    torch.manual_seed(1)
    model = sdnq_post_load_quant(toy_model.build(), **ref["config"])
    _, before = _describe(model)
    if "svd" not in name:       # svd_lowrank draws random test matrices: the factors are reproducible only with the same RNG stream
        assert {k: v["sha1"] for k, v in before.items()} == {k: v["sha1"] for k, v in ref["tensors_before"].items()}
    model = apply_sdnq_options_to_model(model, **ref["options"])
    mods, tensors = _describe(model)
    for mname, e in ref["modules"].items():
        got = mods[mname]
        assert got["class"] == e["class"] and got.get("forward_func") == e.get("forward_func"), mname
        if "dequantizer" in e:
            for k, v in e["dequantizer"].items():
                assert got["dequantizer"].get(k) == v, (mname, k, got["dequantizer"].get(k), v)
    assert sorted(tensors) == sorted(ref["tensors"])
    for key, e in ref["tensors"].items():
        got = tensors[key]
        assert got["dtype"] == e["dtype"] and got["shape"] == e["shape"], key
        assert all(a == b or n == 1 for a, b, n in zip(got["stride"], e["stride"], e["shape"])), (key, got["stride"], e["stride"])
        if "svd" not in name:
            assert got["sha1"] == e["sha1"], f"{key}: bytes differ from the reference's after the flip"


def test_config_to_dict_matches_reference():
    """SDNQConfig(**kwargs).to_dict() -- what `quantization_config.json` of a pre-quantized repo carries (quantizer.py:1075-1079) -- equals
    the reference's for 12 configurations (tests/golden/config_dicts.json, written by the reference's own SDNQConfig).  The skip list
    goes through list(set(...)) on both sides, so its order is compared as a set."""
    from sdnq_b200 import SDNQConfig
    ref = json.load(open(os.path.join(GOLDEN, "config_dicts.json")))
    assert len(ref) >= 12
    for name, e in ref.items():
        d = json.loads(json.dumps(SDNQConfig(**e["kwargs"]).to_dict(), default=str))
        want = dict(e["to_dict"])
        assert sorted(d.pop("modules_to_not_convert")) == sorted(want.pop("modules_to_not_convert")), name
        assert d == want, (name, {k: (d.get(k), want.get(k)) for k in set(d) | set(want) if d.get(k) != want.get(k)})


@pytest.mark.parametrize("name", ["uint4_conv_embedding", "int8_w8a8_conv"])
def test_load_reference_written_checkpoint(name, tmp_path):
    """Checkpoint interchange (SURVEY.md 8 f4 / 8b "config compatibility"): tests/golden/ckpt_<name>/ was written from a model the
    *reference* quantised (state_dict -> safetensors as save_pretrained does, quantization_config.to_dict() -> JSON).  Our loader
    (config -> SDNQ modules with pre_quantized=True -> load_state_dict(assign=True) -> post_process_model) must end up with exactly
    the reference's in-memory model: classes, forward functions, dequantizer metadata, dtypes, shapes, K-major strides, bytes.
    Then our own save -> load round trip must reproduce it again."""
    from sdnq_b200 import load_sdnq_state_dict, save_sdnq_model
    ref = _MODEL_PARITY[name]

    def check(model):
        mods, tensors = _describe(model)
        for mname, e in ref["modules"].items():
            assert mods[mname]["class"] == e["class"] and mods[mname].get("forward_func") == e.get("forward_func"), mname
            for k, v in e.get("dequantizer", {}).items():
                assert mods[mname]["dequantizer"].get(k) == v, (mname, k)
        assert sorted(tensors) == sorted(ref["tensors"])
        for key, e in ref["tensors"].items():
            got = tensors[key]
            assert got["dtype"] == e["dtype"] and got["shape"] == e["shape"] and got["sha1"] == e["sha1"], key
            assert all(a == b or n == 1 for a, b, n in zip(got["stride"], e["stride"], e["shape"])), (key, got["stride"], e["stride"])

    model = load_sdnq_state_dict(toy_model.build(seed=123), os.path.join(GOLDEN, "ckpt_" + name))
    check(model)
    save_sdnq_model(model, str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["model.safetensors", "quantization_config.json"]
    check(load_sdnq_state_dict(toy_model.build(seed=7), str(tmp_path)))
    # load_sdnq_model (reference loader.py:82-196): the model class is built without weights (meta device), the tensors come
    # from the files -- every reader of file_loader.load_files must give the same model
    from sdnq_b200 import load_sdnq_model
    for method in ("safetensors", "threaded"):
        loaded = load_sdnq_model(os.path.join(GOLDEN, "ckpt_" + name), model_cls=toy_model.Toy, model_config={}, load_method=method, dtype=torch.bfloat16)
        check(loaded)
        assert not any(p.is_meta for p in loaded.parameters())
    with pytest.raises(ValueError, match="Unsupported loading method"):
        load_sdnq_model(os.path.join(GOLDEN, "ckpt_" + name), model_cls=toy_model.Toy, model_config={}, load_method="carrier-pigeon")


def test_file_loader_key_mapping_and_shards(tmp_path):
    import torch
    from safetensors.torch import save_file

    from sdnq_b200.file_loader import load_files, map_keys
    save_file({"model.a.weight": torch.arange(4.0), "b": torch.ones(2)}, str(tmp_path / "s1.safetensors"))
    save_file({"model.c.weight": torch.zeros(3)}, str(tmp_path / "s2.safetensors"))
    mapping = {r"^model\.": "", r"weight$": "w"}
    assert map_keys("model.a.weight", mapping) == "a.weight"          # first matching pattern only
    for method in ("safetensors", "threaded"):
        sd = load_files([str(tmp_path / "s1.safetensors"), str(tmp_path / "s2.safetensors")], key_mapping=mapping, method=method)
        assert sorted(sd) == ["a.weight", "b", "c.weight"] and torch.equal(sd["a.weight"], torch.arange(4.0))
    assert sorted(load_files(str(tmp_path / "s2.safetensors"))) == ["model.c.weight"]


@pytest.mark.parametrize("cfg", [dict(weights_dtype="int8", use_quantized_matmul=True), dict(weights_dtype="int4", group_size=32),
                                 dict(weights_dtype="float8_e4m3fn", use_quantized_matmul=True, use_hadamard=True, hadamard_group_size=64)],
                         ids=["w8a8", "dequant_path", "fp8_hadamard"])
@pytest.mark.parametrize("rows", [4, 40])
def test_no_cpu_fallback_anywhere_on_the_forward(cfg, rows):
    """north_star: no CPU / eager fallback.  A CPU tensor raises from every forward (W8A8, dequant path, the rows < 32 branches)
    instead of being computed some other way."""
    import torch

    from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
    from sdnq_b200._lib import SDNQKernelError
    layer, _ = sdnq_quantize_layer(torch.nn.Linear(64, 64).to(torch.bfloat16), SDNQConfig(minimum_allowed_numel=1, **cfg))
    with pytest.raises(SDNQKernelError, match="CUDA"):
        layer(torch.randn(rows, 64, dtype=torch.bfloat16))
    with pytest.raises(SDNQKernelError, match="CUDA"):
        layer.dequantize()


# ------------------------------------------------------------------------------------------------ sibling groups (registration only: no kernels)
def test_sibling_projections_are_found_by_name():
    import torch

    from sdnq_b200 import SDNQConfig, fuse_named_siblings, sdnq_post_load_quant

    class Attention(torch.nn.Module):
        def __init__(self, dim, ctx=None):
            super().__init__()
            self.is_cross_attention = ctx is not None
            self.to_q = torch.nn.Linear(dim, dim)
            self.to_k = torch.nn.Linear(ctx or dim, dim)
            self.to_v = torch.nn.Linear(ctx or dim, dim)
            self.to_out = torch.nn.ModuleList([torch.nn.Linear(dim, dim)])

    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.attn1 = Attention(128)
            self.attn2 = Attention(128, ctx=256)
            self.attn3 = Attention(128, ctx=128)          # cross-attention whose context is as wide as the hidden states

    model = torch.nn.Sequential(Block(), Block()).to(torch.bfloat16)
    model = sdnq_post_load_quant(model, weights_dtype="int8", use_quantized_matmul=True, add_skip_keys=False)
    for blk in model:
        g1 = blk.attn1.to_q.__dict__["_sdnq_siblings"]
        assert [id(m) for m in g1.layers] == [id(blk.attn1.to_q), id(blk.attn1.to_k), id(blk.attn1.to_v)]
        assert "_sdnq_siblings" not in blk.attn2.to_q.__dict__ and "_sdnq_siblings" not in blk.attn3.to_q.__dict__
        assert "_sdnq_siblings" not in blk.attn1.to_out[0].__dict__
    # cross-attention to_k / to_v: pooled over the blocks (every block is handed the same encoder states)
    for name in ("attn2", "attn3"):
        pooled = getattr(model[0], name).to_k.__dict__["_sdnq_siblings"]
        assert [id(m) for m in pooled.layers] == [id(getattr(b, name).__getattr__(p)) for b in model for p in ("to_k", "to_v")]
    # one group per block instead
    from sdnq_b200 import fuse_sibling_projections
    assert fuse_sibling_projections(model, cross_attention_pool=0) == 6
    g2 = model[1].attn2.to_k.__dict__["_sdnq_siblings"]
    assert [id(m) for m in g2.layers] == [id(model[1].attn2.to_k), id(model[1].attn2.to_v)]
    # flat (name, layer) lists: no module attributes, cross-attention is recognised by the input width alone
    named = [(n, m) for n, m in model.named_modules() if m.__class__.__name__ == "SDNQLinear"]
    assert fuse_named_siblings(named, cross_attention_pool=0) == 6
    assert fuse_named_siblings(named) == 2 + 1 + 2          # attn1 x2, pooled attn2 (4 layers); by name and width attn3 looks like self-attention
    pooled = model[0].attn2.to_k.__dict__["_sdnq_siblings"]
    assert len(pooled.layers) == 4 and model[1].attn2.to_v.__dict__["_sdnq_siblings"] is pooled
    # layers on the dequant path form groups of their own kind (one batched library GEMM over their dequantised weights)
    from sdnq_b200.siblings import DequantSiblingGroup, SiblingGroup
    plain = sdnq_post_load_quant(torch.nn.Sequential(Block()).to(torch.bfloat16), weights_dtype="int8", use_quantized_matmul=False, add_skip_keys=False)
    g = plain[0].attn1.to_q.__dict__["_sdnq_siblings"]
    assert type(g) is DequantSiblingGroup and len(g.layers) == 3 and type(model[0].attn1.to_q.__dict__["_sdnq_siblings"]) is SiblingGroup
    # flipping the matmul option re-registers the groups of the new state
    from sdnq_b200 import apply_sdnq_options_to_model
    flipped = apply_sdnq_options_to_model(plain, use_quantized_matmul=True)
    assert type(flipped[0].attn1.to_q.__dict__["_sdnq_siblings"]) is SiblingGroup and g.dead
