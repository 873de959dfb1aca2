"""Dequant path: the weights of the next few layers dequantised together, ahead of their GEMMs.

The reference dequantises inside every forward (layers/linear/forward.py:24-26: SDNQDequantizer.__call__, then F.linear).  At
SD-XL sizes that is 741 dequant launches of 0.4 - 13 M elements per step, each one latency-bound (K3s: ~6 us for 1 - 30 MB).
The weights are frozen and a model calls its layers in the same order every step, so the host layer

  * learns the order in which dequant-path layers ask for their weight (a successor link per layer, re-learned on every call),
  * when a layer asks and nothing is ready, plans a *batch*: this layer and the next ones in the learned order, up to
    SDNQ_B200_PREFETCH_MB of output, dequantised by ONE persistent launch (sdnq_b200_dequant_batch_*) on the side stream into a
    slab of a small ring (no per-layer allocation, a few hundred MB in total -- never a dequantised copy of the model),
  * and, while the GEMMs of that batch run on the caller's stream, launches the following batch.

Every weight is still dequantised once per use by the same kernel arithmetic as the per-layer path (identical bits); a wrong
prediction only leaves a prefetched weight unused.  Safety: a planned batch is tied to the identity of the stored tensors
(forward._StoredState); results are handed out only to the stream / CUDA-graph capture they were produced for; a slab is
rewritten only after the side stream has waited for everything the caller's stream had been given when the batch is launched.
SDNQ_B200_DEQUANT_PREFETCH=0 turns it off (per-layer K3 on the side stream, as before)."""
import os
import weakref

import torch

from . import _lib, ops

MAX_LAYERS = 48
SLABS = 3


def enabled() -> bool:
    return os.environ.get("SDNQ_B200_DEQUANT_PREFETCH", "1") not in ("0", "false", "no")


def _budget_bytes() -> int:
    return max(16, int(os.environ.get("SDNQ_B200_PREFETCH_MB", "256"))) << 20


def eligible(layer, dtype) -> bool:
    """what the batched kernel covers: a Linear with 4-bit integer codes and bf16 SVD factors of rank 16 / 32 / 64, bf16 result"""
    d = layer.sdnq_dequantizer
    up = layer.svd_up
    if (dtype != torch.bfloat16 or up is None or up.dtype != torch.bfloat16 or d.is_conv or d.use_codebook or d.use_hadamard or not d.is_integer
            or d.num_bits != 4 or d.group_size == -2 or layer.weight.ndim > 3 or min(up.shape) not in (16, 32, 64)):
        return False
    N, K = d._linear_nk()
    return K % 32 == 0 and N * K < (1 << 31)


class _Batch:
    __slots__ = ("plan", "entry", "states", "layers", "slab", "event", "tag", "left", "waited", "hits", "streams")


class _Planned:
    """the plans of one chain of layers (one device table per slab of the ring) and the identity of the tensors they embed"""
    __slots__ = ("plans", "states", "good")

    def __init__(self, plans, states):
        self.plans, self.states, self.good = plans, states, False      # good: every weight of the batch was picked up the last time it ran


class Prefetcher:
    """per device"""

    def __init__(self, device):
        self.device = device
        self.succ = {}                 # id(layer) -> weakref of the layer that asked next (same skip flag)
        self.last = None               # (id, capture tag) of the previous request
        self.first = None              # id of the first layer that ever asked: chains do not wrap around the end of a step
        self.slabs = None              # SLABS uint8 device buffers, allocated on first use
        self.slab_busy = [None] * SLABS
        self.plans = {}                # (ids of the chain, skip flag) -> _Planned
        self.ready = {}                # id(layer) -> (_Batch, index)
        self.current = None            # the batch whose weights are being handed out
        self.tag = 0                   # capture id of the previous request

    # ---- bookkeeping ---------------------------------------------------------------------------------------------------------
    def _note(self, layer, tag):
        lid = id(layer)
        if self.first is None:
            self.first = lid
        if self.last is not None and self.last[1] == tag and self.last[0] != lid:
            self.succ[self.last[0]] = weakref.ref(layer)
        self.last = (lid, tag)

    def _chain(self, layer, skip):
        """layer and its learned successors while they are eligible, distinct, and fit the budget"""
        from .forward import _stored_tensors
        chain, seen, nbytes, budget = [], set(), 0, _budget_bytes()
        cur = layer
        while cur is not None and len(chain) < MAX_LAYERS:
            lid = id(cur)
            if lid in seen or (chain and lid == self.first) or lid in self.ready:
                break
            if (not eligible(cur, torch.bfloat16) or bool(cur.sdnq_dequantizer.use_quantized_matmul) != skip
                    or any(t is not None and t.device != self.device for t in _stored_tensors(cur))):
                break
            N, K = cur.sdnq_dequantizer._linear_nk()
            size = (N * K * 2 + 255) // 256 * 256
            if chain and nbytes + size > budget:
                break
            if size > budget:
                break
            chain.append(cur)
            seen.add(lid)
            nbytes += size
            ref = self.succ.get(lid)
            cur = ref() if ref is not None else None
        return chain

    def _free_slab(self):
        for k in range(SLABS):
            b = self.slab_busy[k]
            if b is None or b.left == 0:
                return k
        for k in range(SLABS):                                # nobody came for the rest of an old batch (the order changed): abandon it
            b = self.slab_busy[k]
            if b is not self.current and b.waited:
                return k
        return None

    def _drop(self, b):
        for ref in b.layers:
            layer = ref()
            if layer is not None and self.ready.get(id(layer), (None,))[0] is b:
                del self.ready[id(layer)]

    def _launch(self, chain, skip, tag, main, side, ahead=False):
        """plan (or reuse the plans of) `chain` and launch it into a free slab on the side stream; None if that is not possible now"""
        from .forward import _StoredState, _stored_tensors
        key = (tuple(id(layer) for layer in chain), skip)
        entry = self.plans.get(key)
        if entry is not None and not all(st.matches(layer, _stored_tensors(layer)) for st, layer in zip(entry.states, chain)):
            entry = None
            del self.plans[key]
        if entry is None:
            if tag != 0:
                return None                                  # planning uploads tables: never inside a graph capture
            if self.slabs is None:
                self.slabs = [torch.empty(_budget_bytes(), dtype=torch.uint8, device=self.device) for _ in range(SLABS)]
            jobs = []
            for layer in chain:
                d = layer.sdnq_dequantizer
                N, K = d._linear_nk()
                jobs.append(dict(weight=layer.weight, weights_dtype=d.weights_dtype, scale=layer.scale, zero_point=layer.zero_point, N=N, K=K,
                                 group_size=d.group_size, svd_up=layer.svd_up, svd_down=layer.svd_down, svd_layout_matmul=skip))
            try:
                plans = [ops.dequant_batch_plan(jobs, slab) for slab in self.slabs]
            except _lib.SDNQKernelError:
                return None
            entry = _Planned(plans, [_StoredState(layer, _stored_tensors(layer)) for layer in chain])
            if len(self.plans) > 4096:
                self.plans.clear()
            self.plans[key] = entry
        if ahead and tag != 0 and not entry.good:
            return None          # inside a capture a batch launched ahead must be joined by its consumers: only chains that were fully used before
        k = self._free_slab()
        if k is None:
            return None
        b = _Batch()
        b.plan, b.entry, b.states, b.layers, b.slab, b.tag = entry.plans[k], entry, entry.states, [weakref.ref(layer) for layer in chain], k, tag
        b.left, b.waited, b.hits, b.streams = len(chain), False, 0, {}
        # the slab's previous readers (GEMMs already handed to the caller's stream -- and to any other stream its weights were handed
        # out on) finish before it is rewritten; the weights were written by whatever the caller's stream did before as well
        side.wait_stream(main)
        old = self.slab_busy[k]
        if old is not None:
            for st in old.streams.values():
                if st.cuda_stream != main.cuda_stream:
                    side.wait_stream(st)
        with torch.cuda.stream(side):
            ops.dequant_batch_run(b.plan)
            b.event = torch.cuda.Event()
            b.event.record(side)
        if old is not None:                                   # entries of an abandoned batch must not be served from a rewritten slab
            self._drop(old)
        self.slab_busy[k] = b
        for i, layer in enumerate(chain):
            self.ready[id(layer)] = (b, i)
        return b

    # ---- the request -----------------------------------------------------------------------------------------------------------
    def get(self, layer, skip: bool, main, side):
        """the dequantised weight of `layer` (bf16 [N,K], visible to `main`), or None: the caller dequantises it itself"""
        from .forward import _stored_tensors
        tag = ops.capture_id(main.cuda_stream)
        if tag != self.tag:            # a graph capture began or ended: nothing produced on the other side of it is handed out
            self.ready.clear()
            self.slab_busy = [None] * SLABS
            self.current, self.last, self.tag = None, None, tag
        self._note(layer, tag)
        lid = id(layer)
        hit = self.ready.pop(lid, None)
        if hit is not None:
            b, i = hit
            b.left -= 1
            if not b.states[i].matches(layer, _stored_tensors(layer)):
                hit = None                                   # the stored tensors were replaced since the batch was planned
            else:
                b.hits += 1
                if b.left == 0 and b.hits == len(b.layers):
                    b.entry.good = True
        if hit is None:
            chain = self._chain(layer, skip)
            if len(chain) < 2:
                return None
            b = self._launch(chain, skip, tag, main, side)
            if b is None:
                return None
            self.ready.pop(lid, None)
            b.left -= 1
            b.hits += 1
            i = 0
        first_of_batch = not b.waited
        if main.cuda_stream not in b.streams:
            main.wait_event(b.event)                         # once per batch and stream: everything in it is visible to that stream from here on
            b.streams[main.cuda_stream] = main
            b.waited = True
            self.current = b
        W = b.plan.outs[i]
        if first_of_batch:
            # while this batch's GEMMs run: the batch that follows it in the learned order
            tail = b.layers[-1]()
            ref = self.succ.get(id(tail)) if tail is not None else None
            nxt = ref() if ref is not None else None
            if nxt is not None and id(nxt) != self.first and id(nxt) not in self.ready:
                chain = self._chain(nxt, skip)
                if len(chain) >= 2:
                    self._launch(chain, skip, tag, main, side, ahead=True)
        return W


_PREFETCHERS: dict = {}


def prefetcher(device) -> Prefetcher:
    p = _PREFETCHERS.get(device.index)
    if p is None:
        p = _PREFETCHERS[device.index] = Prefetcher(device)
    return p


def reset():
    """forget everything learned (tests; after a model was unloaded)"""
    _PREFETCHERS.clear()
