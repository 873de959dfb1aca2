"""Sibling projections as one launch.

to_q / to_k / to_v of a self-attention block, the to_k / to_v pair of a cross-attention block, add_{q,k,v}_proj, gate / up
projections: Linears that are handed the very same activation tensor.  The reference runs each of them as its own forward
(layers/linear/linear_int8.py:100-125: quantise the input again, one scaled matmul); at SD-XL bs=1 sizes every one of those
launches is a one-wave GEMM on half of the SMs, dominated by launch and pipeline-fill latency.

Here the siblings of a group keep their matmul operands in ONE buffer (their [N_g, K] operands stacked along N, every segment
padded to the tile width); the first sibling that is called with an activation tensor runs K2 once and ONE grouped K1 launch
(sdnq_b200_scaled_mm_grouped) that writes every sibling's own contiguous output, and the other siblings pick their result up when
they are called with the same tensor (same storage, shape, strides, dtype and in-place version -- the rule the K2 cache uses; a
strong reference to the input is held meanwhile).  Every output element is exactly what the un-grouped forward computes: same
codes, same accumulator, same epilogue.  A sibling that is called with a different tensor simply starts a new grouped launch for
that tensor, so a wrong guess about who shares inputs costs time, never correctness.

Groups are registered by `fuse_sibling_projections(model)` (called by sdnq_post_load_quant and the loader; module names as
Diffusers / Transformers use them) or explicitly with `group_siblings([layer, ...])`.  SDNQ_B200_SIBLINGS=0 turns grouping off."""
import os
import weakref

import torch

from . import ops
from .common import dtype_dict

MAX_GROUP = 8            # siblings per launch (output tensor maps in the kernel's parameter block)
_WASTE_LIMIT = 3         # consecutive launches with mostly unused outputs before a group is dissolved
_REBUILD_LIMIT = 4       # a group whose stored tensors keep changing (CPU offload re-uploads them every step) is dissolved

# children of one parent module that are fed the same tensor
SIBLING_NAME_SETS = (
    ("to_q", "to_k", "to_v"), ("add_q_proj", "add_k_proj", "add_v_proj"), ("q_proj", "k_proj", "v_proj"), ("query", "key", "value"),
    ("gate_proj", "up_proj"), ("w1", "w3"), ("to_k", "to_v"), ("k_proj", "v_proj"), ("add_k_proj", "add_v_proj"),
)


def siblings_enabled() -> bool:
    return os.environ.get("SDNQ_B200_SIBLINGS", "1") not in ("0", "false", "no")


def _kind(layer):
    """"mm": a Linear on the quantized-matmul path the grouped K1 launch covers; "dq": a Linear on the dequant path; None: neither"""
    d = getattr(layer, "sdnq_dequantizer", None)
    if d is None or layer.__class__.__name__ != "SDNQLinear" or d.is_conv or len(tuple(d.original_shape)) != 2:
        return None
    if not d.use_quantized_matmul:
        return "dq"
    if getattr(layer, "svd_up", None) is not None or (layer.weight.ndim != 2 and not d.is_packed):
        return None
    return "mm" if d.quantized_matmul_dtype in ("int8", "uint8", "float8_e4m3fn") else None


def _eligible(layer) -> bool:
    return _kind(layer) == "mm"


def _signature(layer):
    d = layer.sdnq_dequantizer
    N, K = d.matmul_nk() if d.use_quantized_matmul else tuple(d.original_shape)
    if not d.use_quantized_matmul:
        return ("dq", N, K)                  # one batched library GEMM: the members' weights have one shape
    return (K, d.quantized_matmul_dtype, d.hadamard_group_size if d.use_hadamard else 0)


class SiblingGroup:
    def __init__(self, layers):
        self.layers = list(layers)
        self.index = {id(layer): i for i, layer in enumerate(self.layers)}
        self.state = None            # (per-layer matmul operands the buffers were built from, buffers...)
        self.cache = None            # (key, input kept alive, outputs not yet handed out)
        self.rebuilds = 0
        self.wasted = 0              # consecutive launches at least half of whose outputs were never picked up
        self.dead = False

    # ---- the concatenated operand -------------------------------------------------------------------------------------------
    def _current(self, ops_now) -> bool:
        st = self.state
        return st is not None and all(a is b for a, b in zip(st["ops"], ops_now))

    def _build(self, ops_now):
        from .forward import _StoredState
        first = ops_now[0]
        packed = first.packed
        if any(o.packed != packed or (o.zp is None) != (first.zp is None) or (o.colsum is None) != (first.colsum is None)
               or o.wq.dtype != first.wq.dtype or o.wq.device != first.wq.device for o in ops_now):
            return False
        K = self.layers[0].sdnq_dequantizer.matmul_nk()[1]
        ns = [layer.sdnq_dequantizer.matmul_nk()[0] for layer in self.layers]
        if any(n % 8 for n in ns):
            return False
        align = 256 if all(n % 256 == 0 for n in ns) and packed is None else 128
        starts = [0]
        for n in ns:
            starts.append(starts[-1] + (n + align - 1) // align * align)
        dev = first.wq.device
        row_bytes = K * dtype_dict[packed]["num_bits"] // 8 if packed is not None else K
        wq_cat = torch.zeros((starts[-1], row_bytes), dtype=first.wq.dtype, device=dev)

        def cat_vec(vals, dtype):
            out = torch.zeros(starts[-1], dtype=dtype, device=dev)
            for v, s, n in zip(vals, starts, ns):
                if v is not None:
                    out[s:s + n] = v.reshape(-1).to(dtype)
            return out
        for o, s, n in zip(ops_now, starts, ns):
            wq_cat[s:s + n] = o.wq.reshape(n, row_bytes)
        biases = [layer.bias for layer in self.layers]
        state = {
            "ops": list(ops_now), "wq": wq_cat, "sw": cat_vec([o.sw for o in ops_now], torch.float32),
            "zp": cat_vec([o.zp for o in ops_now], torch.float32) if first.zp is not None else None,
            "colsum": cat_vec([o.colsum for o in ops_now], torch.int32) if first.colsum is not None else None,
            "bias": cat_vec(biases, torch.float32) if any(b is not None for b in biases) else None,
            "bias_marks": [None if b is None else (weakref.ref(b), b.data_ptr(), b._version if not b.is_inference() else 0) for b in biases],
            "starts": starts, "ns": ns, "packed": packed,
        }
        # derived per-layer operands become views of the group's buffer (one copy of those codes in memory); a stored weight that is
        # itself the operand (row-wise 8-bit, K-major) stays where it is unless SDNQ_B200_SIBLINGS_INPLACE=1
        for layer, o, s, n in zip(self.layers, ops_now, starts, ns):
            view = wq_cat[s:s + n]
            w = layer.weight
            k_major = w.ndim == 2 and not w.is_contiguous() and w.t().is_contiguous()
            shares = (w.dtype == view.dtype and w.numel() == view.numel() and (w.is_contiguous() or k_major)
                      and w.data_ptr() == o.wq.data_ptr())
            if not shares:
                o.wq = view                   # a derived operand (re-quantised / re-centred copy): keep only the group's copy
            elif os.environ.get("SDNQ_B200_SIBLINGS_INPLACE", "0") not in ("0", "false", "no", ""):
                # opt-in: the stored weight itself moves into the group's buffer (no second copy of the codes; the siblings' weights
                # then share one storage, which checkpoint writers that look for tied tensors will notice)
                o.wq = view
                w.data = view.t() if k_major else view.reshape(w.shape)
                o.key = (_StoredState(layer, (layer.weight, layer.scale, layer.zero_point)), o.key[1])
        self.state = state
        return True

    def _bias_current(self) -> bool:
        for layer, mark in zip(self.layers, self.state["bias_marks"]):
            b = layer.bias
            if (b is None) != (mark is None):
                return False
            if b is not None and (mark[0]() is not b or mark[1] != b.data_ptr() or mark[2] != (b._version if not b.is_inference() else 0)):
                return False
        return True

    # ---- forward -------------------------------------------------------------------------------------------------------------
    def forward(self, layer, x2: torch.Tensor, out_dtype: torch.dtype):
        """The output [M, N] of `layer` for the activations x2, or None when the group cannot serve it (the caller then runs the
        layer's own launch)."""
        from .forward import _version, matmul_operand, quantized_activations
        if self.dead or not x2.is_cuda:
            return None
        idx = self.index[id(layer)]
        stream = torch.cuda.current_stream(x2.device).cuda_stream
        key = (x2.data_ptr(), _version(x2), tuple(x2.shape), tuple(x2.stride()), x2.dtype, out_dtype, x2.device.index, stream, ops.capture_id(stream))
        c = self.cache
        if c is not None and c[0] == key and c[2][idx] is not None:
            y, c[2][idx] = c[2][idx], None
            if all(o is None for o in c[2]):
                self.cache = None
            return y
        if not all(_eligible(sib) for sib in self.layers) or len({_signature(sib) for sib in self.layers}) != 1:
            self.dissolve()                  # a sibling was switched to another path (apply_sdnq_options_to_model) or replaced
            return None
        ops_now = [matmul_operand(sib) for sib in self.layers]
        if not self._current(ops_now) or not self._bias_current():
            self.rebuilds += 1
            if self.rebuilds > _REBUILD_LIMIT or not self._build(ops_now):
                self.dissolve()
                return None
        st = self.state
        d = layer.sdnq_dequantizer
        mm = d.quantized_matmul_dtype
        hg = d.hadamard_group_size if d.use_hadamard else 0
        xq, sx, zx, rowsum, _ = quantized_activations(x2, mm, hg, st["zp"] is not None, False)
        outs = ops.scaled_mm_grouped(xq, st["wq"], sx, st["sw"], st["starts"], st["ns"], st["bias"], out_dtype, rowsum=rowsum, zp=st["zp"],
                                     colsum=st["colsum"], zx=zx, packed_dtype=st["packed"])
        y, outs[idx] = outs[idx], None
        if c is not None:
            # results of the previous launch that nobody asked for: the siblings are not fed the same tensor after all.  A group that
            # keeps computing outputs for nothing is dissolved (its members go back to their own launches).
            unused = sum(o is not None for o in c[2])
            self.wasted = self.wasted + 1 if 2 * unused >= len(self.layers) else 0
            if self.wasted > _WASTE_LIMIT:
                self.dissolve()
                return y
        self.cache = (key, x2, outs)
        return y

    def dissolve(self):
        self.dead = True
        self.state = self.cache = None
        for layer in self.layers:
            if layer.__dict__.get("_sdnq_siblings") is self:
                del layer.__dict__["_sdnq_siblings"]


class DequantSiblingGroup(SiblingGroup):
    """Siblings on the dequant path (use_quantized_matmul=False): the reference runs F.linear(x, dequantised weight, bias) per layer
    (layers/linear/forward.py:24-26).  When the members' dequantised weights sit next to each other in memory -- they do when the
    batched dequantiser (prefetch.py) produced them: consecutive layers of the learned call order are consecutive in its slab -- the
    members' GEMMs run as ONE strided-batched library GEMM over that slab (weights [G, N, K], activations broadcast, outputs
    [G, M, N]: every member's output is its own contiguous slice).  Otherwise every member runs its own F.linear, as before."""

    def _stacked_bias(self, biases, N, like):
        """[G, 1, N] biases (zeros for members without one) in the GEMM's dtype, rebuilt when a member's bias changes"""
        marks = tuple(None if b is None else (id(b), b.data_ptr(), b._version if not b.is_inference() else 0) for b in biases) + (like.dtype, like.device)
        if self.state is None or self.state[0] != marks:
            stacked = torch.stack([torch.zeros(N, dtype=like.dtype, device=like.device) if b is None else b.detach().to(like.dtype) for b in biases]).unsqueeze(1)
            self.state = (marks, stacked, list(biases))          # (the bias tensors are kept alive so that their ids stay theirs)
        return self.state[1]

    def forward(self, layer, input: torch.Tensor):
        from .forward import _dequant_weight_overlapped, _version
        if self.dead or not input.is_cuda:
            return None
        idx = self.index[id(layer)]
        K = input.shape[-1]
        x2 = input.reshape(-1, K)
        stream = torch.cuda.current_stream(input.device).cuda_stream
        key = (x2.data_ptr(), _version(input), tuple(x2.shape), tuple(x2.stride()), x2.dtype, input.device.index, stream, ops.capture_id(stream))
        c = self.cache
        if c is not None and c[0] == key and c[2][idx] is not None:
            y, c[2][idx] = c[2][idx], None
            if all(o is None for o in c[2]):
                self.cache = None
            return y
        if any(_kind(sib) != "dq" for sib in self.layers) or len({_signature(sib) for sib in self.layers}) != 1:
            self.dissolve()
            return None
        weights = [_dequant_weight_overlapped(sib, input, False) for sib in self.layers]
        w0 = weights[0]
        G, (N, Kw) = len(weights), w0.shape
        step = weights[1].data_ptr() - w0.data_ptr() if G > 1 else 0
        adjacent = (Kw == K and step > 0 and step % w0.element_size() == 0 and w0.is_contiguous()
                    and all(w.shape == w0.shape and w.dtype == w0.dtype and w.is_contiguous() and w.untyped_storage().data_ptr() == w0.untyped_storage().data_ptr()
                            and w.data_ptr() - w0.data_ptr() == i * step for i, w in enumerate(weights)))
        biases = [sib.bias for sib in self.layers]
        if adjacent:
            w3 = w0.as_strided((G, N, K), (step // w0.element_size(), K, 1))
            xb = x2.unsqueeze(0).expand(G, -1, -1)
            if all(b is None for b in biases):
                out3 = torch.bmm(xb, w3.transpose(1, 2))
            else:
                out3 = torch.baddbmm(self._stacked_bias(biases, N, w0), xb, w3.transpose(1, 2))
            outs = [out3[i] for i in range(G)]
        else:
            outs = [torch.nn.functional.linear(x2, w, b) for w, b in zip(weights, biases)]
        y, outs[idx] = outs[idx], None
        if c is not None:
            unused = sum(o is not None for o in c[2])
            self.wasted = self.wasted + 1 if 2 * unused >= len(self.layers) else 0
            if self.wasted > _WASTE_LIMIT:
                self.dissolve()
                return y
        self.cache = (key, input, outs)
        return y


def group_siblings(layers) -> SiblingGroup | None:
    """Register `layers` (SDNQLinear modules that are fed the same tensor) as one group: Linears on the quantized-matmul path share one
    grouped K1 launch, Linears on the dequant path one batched library GEMM over their dequantised weights.  Returns None when they
    cannot be grouped (different K / matmul dtype / rotation, SVD layers on the matmul path, convolutions, mixed paths, fewer than two)."""
    kinds = {_kind(layer) for layer in layers}
    if len(kinds) != 1 or None in kinds:
        return None
    layers = list(layers)
    if len(layers) < 2 or len(layers) > MAX_GROUP or len({_signature(layer) for layer in layers}) != 1 or len({id(layer) for layer in layers}) != len(layers):
        return None
    for layer in layers:
        old = layer.__dict__.get("_sdnq_siblings")
        if old is not None:
            old.dissolve()
    group = SiblingGroup(layers) if kinds == {"mm"} else DequantSiblingGroup(layers)
    for layer in layers:
        layer.__dict__["_sdnq_siblings"] = group
    return group


def _family_groups(children: dict):
    """children: name -> module of one parent.  Yields lists of modules that share an input."""
    used = set()
    for names in SIBLING_NAME_SETS:
        if all(n in children for n in names) and not any(n in used for n in names):
            mods = [children[n] for n in names]
            if _kind(mods[0]) is None or any(_kind(m) != _kind(mods[0]) for m in mods):
                continue
            if len({_signature(m) for m in mods}) != 1:
                continue       # cross-attention: to_q reads the hidden states, to_k / to_v the encoder states (a later, smaller name set matches)
            used.update(names)
            yield mods


def _register(families, cross_attention_pool: int) -> int:
    """families: lists of modules that share an input, in model order.  to_k / to_v style pairs whose input width differs from what
    their block's query projection reads -- cross-attention: every such block of a UNet / DiT is handed the same
    encoder_hidden_states tensor -- are pooled over up to cross_attention_pool layers of *different* blocks."""
    count, pools = 0, {}
    for mods, cross in families:
        if cross and cross_attention_pool >= 4:
            pools.setdefault(_signature(mods[0]), []).append(mods)
        elif group_siblings(mods) is not None:
            count += 1
    per_group = max(1, min(cross_attention_pool, MAX_GROUP) // 2)
    for pairs in pools.values():
        for i in range(0, len(pairs), per_group):
            if group_siblings([m for pair in pairs[i:i + per_group] for m in pair]) is not None:
                count += 1
    return count


def _families(children: dict, cross_flag: bool):
    """(modules, is_cross_attention_pair) for one parent's children."""
    if cross_flag:
        children = {k: v for k, v in children.items() if k not in ("to_q", "q_proj", "query")}
    for mods in _family_groups(children):
        q = next((children[n] for n in ("to_q", "q_proj", "query") if n in children and _kind(children[n]) is not None), None)
        q_in = None if q is None else q.sdnq_dequantizer.original_shape[1]
        cross = len(mods) == 2 and (cross_flag or (q is not None and q not in mods and q_in != mods[0].sdnq_dequantizer.original_shape[1]))
        yield mods, cross


def fuse_sibling_projections(model: torch.nn.Module, cross_attention_pool: int = MAX_GROUP) -> int:
    """Find sibling projections by their Diffusers / Transformers names under every parent module and register them as groups.
    A cross-attention block whose to_q has the same input width as its to_k / to_v cannot be told apart from self-attention by
    shape: it is recognised by the parent's `is_cross_attention` attribute.  Cross-attention to_k / to_v pairs are pooled over
    cross_attention_pool layers (0: one group per block).  Wrong guesses cost time only, and a group whose results keep going
    unused dissolves itself."""
    fams = []
    for parent in model.modules():
        fams += list(_families(dict(parent.named_children()), bool(getattr(parent, "is_cross_attention", False))))
    return _register(fams, cross_attention_pool)


def fuse_named_siblings(named_layers, cross_attention_pool: int = MAX_GROUP) -> int:
    """The same for a flat list of (qualified name, layer) pairs (no module attributes to look at: cross-attention is recognised
    by the input width of to_k / to_v differing from to_q's)."""
    parents: dict = {}
    for name, layer in named_layers:
        parent, _, leaf = name.rpartition(".")
        parents.setdefault(parent, {})[leaf] = layer
    fams = []
    for children in parents.values():
        fams += list(_families(children, False))
    return _register(fams, cross_attention_pool)
