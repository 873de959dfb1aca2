"""Benchmark of the quantized-Linear hot path (contract: see the repo brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sdxl_int8|flux_fp8] [--impl reference]

One "step" = one pass over every SDNQ-quantised Linear of one denoise step of the named model on synthetic activations of
the real shapes, each layer with its own random weights quantised through the public API
(`sdnq_quantize_layer(nn.Linear, SDNQConfig)`), run through `SDNQLinear.forward` -> C ABI -> sm_100a kernels.

  value  whole-job TFLOP/s (2*M*N*K over all Linears of the step), inputs resident in HBM, step replayed as a CUDA graph
  e2e    same metric through the same call with HOST inputs: pinned host -> device copy of the step inputs and a device ->
         host read of the step output inside the timed region
  roofline      the dominant kernel (tcgen05 W8A8 GEMM): algorithmic FLOPs of all its launches in a step / their CUDA-event time
                (the launches replayed back to back as a graph, activations L2-hot and weights cold exactly as in the step)
  cpu_baseline  the oracle port (numpy) on the host cores, bounded sample of the same workload

N > 1 (torchrun): every rank runs the same stack on its own batch shard (weights replicated, no data-path collective), one
NCCL all_gather of the step output per step; time = max over ranks; scaling = weak.
`--impl reference`: the reference's CPU algorithm (oracle port) on the host cores, rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------ workloads
def sdxl_linears():
    """SD-XL UNet, bs = 1, 1024x1024 (latents 128^2): 4096 tokens at C = 640 (10 transformer layers in 5 Transformer2DModels),
    1024 tokens at C = 1280 (60 layers in 6 models), 77 text tokens of width 2048; plus the M = batch Linears the reference also
    quantises (ResNet time_emb_proj, add_embedding) which take the small-M dequant path.  -> list of (name, M, N, K)."""
    layers = []
    for C, M, models, per_model in ((640, 4096, 5, 2), (1280, 1024, 6, 10)):
        for t in range(models):
            layers.append((f"c{C}.t{t}.proj_in", M, C, C))
            for blk in range(per_model):
                p = f"c{C}.t{t}.b{blk}"
                layers += [(f"{p}.attn1.to_q", M, C, C), (f"{p}.attn1.to_k", M, C, C), (f"{p}.attn1.to_v", M, C, C), (f"{p}.attn1.to_out", M, C, C),
                           (f"{p}.attn2.to_q", M, C, C), (f"{p}.attn2.to_k", 77, C, 2048), (f"{p}.attn2.to_v", 77, C, 2048),
                           (f"{p}.attn2.to_out", M, C, C), (f"{p}.ff.proj", M, 8 * C, C), (f"{p}.ff.out", M, C, 4 * C)]
            layers.append((f"c{C}.t{t}.proj_out", M, C, C))
    for i, c_out in enumerate([320] * 2 + [640] * 2 + [1280] * 2 + [1280] * 2 + [1280] * 3 + [640] * 3 + [320] * 3):
        layers.append((f"resnet{i}.time_emb_proj", 1, c_out, 1280))
    layers += [("add_embedding.linear_1", 1, 1280, 2816), ("add_embedding.linear_2", 1, 1280, 1280)]
    return layers


def flux_linears(batch=4):
    """FLUX.1-dev DiT, 1024x1024, bs = `batch`: 4096 image tokens + 512 text tokens per image, D = 3072; 19 double blocks
    (separate image / text streams), 38 single blocks; AdaLN Linears have M = batch (small-M path)."""
    D, layers = 3072, []
    mi, mt = 4096 * batch, 512 * batch
    for b in range(19):
        for stream, M in (("img", mi), ("txt", mt)):
            p = f"double{b}.{stream}"
            layers += [(f"{p}.norm1.linear", batch, 6 * D, D), (f"{p}.to_q", M, D, D), (f"{p}.to_k", M, D, D), (f"{p}.to_v", M, D, D),
                       (f"{p}.to_out", M, D, D), (f"{p}.ff.proj", M, 4 * D, D), (f"{p}.ff.out", M, D, 4 * D)]
    for b in range(38):
        p, M = f"single{b}", mi + mt
        if b > 0:
            layers.append((f"{p}.norm.linear", batch, 3 * D, D))
        layers += [(f"{p}.to_q", M, D, D), (f"{p}.to_k", M, D, D), (f"{p}.to_v", M, D, D), (f"{p}.proj_mlp", M, 4 * D, D),
                   (f"{p}.proj_out", M, D, 5 * D)]
    return layers


WORKLOADS = {
    "sdxl_int8": dict(describe="SD-XL UNet int8 W8A8 (use_quantized_matmul=True), bs=1, 1024x1024: all quantised Linears of one denoise step",
                      layers=sdxl_linears, config=dict(weights_dtype="int8", use_quantized_matmul=True), dtype="int8"),
    "sdxl_int4_svd_dequant": dict(describe="SD-XL UNet int4 group_size=128 + SVD rank 32, dequant-to-bf16-GEMM path (use_quantized_matmul=False), bs=1, 1024x1024",
                                  layers=sdxl_linears, config=dict(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32), dtype="bf16"),
    "flux_fp8": dict(describe="FLUX.1-dev DiT float8_e4m3fn + Hadamard(256) W8A8, bs=4, 1024x1024: all quantised Linears of one denoise step",
                     layers=flux_linears, config=dict(weights_dtype="float8_e4m3fn", use_quantized_matmul=True, use_hadamard=True, hadamard_group_size=256),
                     dtype="f8e4m3"),
}


def total_flops(layers):
    return sum(2.0 * m * n * k for _, m, n, k in layers)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})

        def num(v):
            try:
                return float(v)
            except ValueError:
                return float("nan")
        return {"sm_mhz": statistics.median(num(r[0]) for r in rows), "sm_max_mhz": num(rows[0][1]),
                "power_w_max": max(num(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------ CPU oracle arm
def cpu_oracle_sample(workload, budget_s=15.0, seed=0):
    """Time the numpy oracle (reference algorithm restated) on a bounded sample: one layer of each distinct (N, K) with M capped
    at 256 rows; throughput is FLOPs of what was actually computed / wall time."""
    import numpy as np

    from oracle import sdnq_oracle as O
    spec = WORKLOADS[workload]
    rng = np.random.default_rng(seed)
    shapes, seen = [], set()
    for _, m, n, k in spec["layers"]():
        if (n, k) not in seen and m >= 32:
            seen.add((n, k))
            shapes.append((min(m, 256), n, k))
    fp8 = spec["config"]["weights_dtype"].startswith("float8")
    hg = spec["config"].get("hadamard_group_size", 256) if spec["config"].get("use_hadamard") else 0
    prepared = []
    dequant_path = not spec["config"].get("use_quantized_matmul", False)
    for m, n, k in shapes:
        w = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
        if dequant_path:       # int4 g128 + SVD rank 32: packed codes, group scales, low-rank factors; dequant then f32 GEMM
            gs, r = spec["config"]["group_size"], spec["config"]["svd_rank"]
            wg = w.reshape(n, k // gs, gs)
            sc = (np.abs(wg).max(axis=-1, keepdims=True) / 7).astype(np.float32)
            codes = np.clip(np.rint(wg / sc), -8, 7).astype(np.int64)
            layer = O.Layer(O.pack_int(codes, "int4"), sc, None, O.bf16_round(rng.standard_normal((n, r)).astype(np.float32) * 0.05),
                            O.bf16_round(rng.standard_normal((r, k)).astype(np.float32) * 0.05),
                            bias=O.bf16_round(rng.standard_normal(n).astype(np.float32)), weights_dtype="int4", quantized_weight_shape=[n, k // gs, gs],
                            result_shape=[n, k], group_size=gs, use_quantized_matmul=False)
            x = O.bf16_round(rng.standard_normal((m, k)).astype(np.float32))
            prepared.append((layer, x, 2.0 * m * n * k))
            continue
        if fp8:
            wq, sw = O.quantize_fp_mm(w, axis=-1)
        else:
            wq, sw = O.quantize_int_mm(w, axis=-1)
        meta = dict(weights_dtype=spec["config"]["weights_dtype"], quantized_matmul_dtype="float8_e4m3fn" if fp8 else "int8",
                    quantized_weight_shape=[k, n], group_size=-1, use_quantized_matmul=True, use_hadamard=bool(hg),
                    hadamard_group_size=hg if hg and k % hg == 0 else 128)
        layer = O.Layer(np.ascontiguousarray(wq.T), np.ascontiguousarray(sw.T), bias=O.bf16_round(rng.standard_normal(n).astype(np.float32)), **meta)
        x = O.bf16_round(rng.standard_normal((m, k)).astype(np.float32))
        prepared.append((layer, x, 2.0 * m * n * k))
    flops, t0, reps = 0.0, time.perf_counter(), 0
    while True:
        for layer, x, fl in prepared:
            O.linear_forward(layer, x)
            flops += fl
        reps += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": flops / dt / 1e12, "unit": "TFLOP/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{len(prepared)} layers (one per distinct NxK of the workload, M capped at 256 rows) x {reps} passes, numpy oracle port of the reference CPU-eager path, {dt:.1f} s"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = WORKLOADS[args.workload]
    t0 = time.perf_counter()
    per_step = max(2.0, min(20.0, 90.0 / max(args.steps + args.warmup, 1)))
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_oracle_sample(args.workload, budget_s=per_step, seed=i)
        if i >= args.warmup:
            vals.append(r)
    value = statistics.mean(v["value"] for v in vals)
    wall = time.perf_counter() - t0
    line = {"impl": "reference", "metric": "quantized_linear_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps + args.warmup, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": spec["dtype"], "data": "synthetic",
            "config": {"workload": args.workload, "describe": spec["describe"], "device": "host CPU"},
            "cpu_baseline": dict(vals[-1], value=value),
            "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def build_stack(workload, device):
    import torch

    from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
    spec = WORKLOADS[workload]
    cfg = spec["config"]
    stack = []
    torch.manual_seed(1234)
    for name, m, n, k in spec["layers"]():
        lin = torch.nn.Linear(k, n, bias=True, device=device, dtype=torch.bfloat16)
        layer, _ = sdnq_quantize_layer(lin, SDNQConfig(**cfg), param_name=name + ".weight")
        stack.append((name, m, n, k, layer))
    return stack


def make_inputs(stack, device, host_inputs):
    """activation buffers of the step, all derived on device from the step inputs (`host_inputs`: name -> pinned host tensor)."""
    import torch
    dev_in = {k: torch.empty_like(v, device=device) for k, v in host_inputs.items()}
    return dev_in


def derive_activations(dev_in, shapes):
    """(M,K) activations for every distinct input shape of the stack, cut out of the step inputs without extra HBM-resident
    copies where possible (views) -- stand-ins for the attention / norm outputs that feed the Linears in the real model."""
    import torch
    base = dev_in["hidden"]
    flat = base.reshape(-1)
    acts = {}
    for (m, k) in shapes:
        need = m * k
        if need <= flat.numel():
            acts[(m, k)] = flat[:need].view(m, k)
        else:
            reps = (need + flat.numel() - 1) // flat.numel()
            acts[(m, k)] = flat.repeat(reps)[:need].view(m, k)
    if "context" in dev_in:
        ctx = dev_in["context"]
        for (m, k) in shapes:
            if (m, k) == tuple(ctx.shape):
                acts[(m, k)] = ctx
    return acts


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from sdnq_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the one JSON line (the version banner goes to stdout)
        dist.init_process_group("nccl", device_id=device)
    _lib.check(_lib.load().sdnq_b200_check_device(local_rank))
    spec = WORKLOADS[args.workload]
    stack = build_stack(args.workload, device)
    layers_meta = [(n, m, nn_, k) for n, m, nn_, k, _ in stack]
    flops_step = total_flops(layers_meta)

    # ---- step inputs live on the host (pinned); everything else is derived on device
    torch.manual_seed(100 + rank)
    hidden_shape = next((m, k) for _, m, _, k in layers_meta if m >= 32)      # the first Linear's input: the step's "latents"
    host_inputs = {"hidden": torch.randn(hidden_shape, dtype=torch.bfloat16).pin_memory()}
    ctx_shapes = [(m, k) for _, m, _, k in layers_meta if m == 77]
    if ctx_shapes:
        host_inputs["context"] = torch.randn(ctx_shapes[0], dtype=torch.bfloat16).pin_memory()
    dev_in = {k: v.to(device, non_blocking=True) for k, v in host_inputs.items()}
    shapes = sorted({(m, k) for _, m, _, k in layers_meta})

    out_holder = {}

    def step():
        acts = derive_activations(dev_in, shapes)
        y = None
        for _, m, _, k, layer in stack:
            y = layer(acts[(m, k)])
            if m >= 32:
                out_holder["y"] = y
        return out_holder["y"]

    # ---- warm-up (eager): builds operand caches, sets kernel attributes, sizes workspaces
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    _lib.launch_count(reset=True)
    step()
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count()

    # ---- capture the whole step in one CUDA graph (launch-bound inner loop: ~1.5k launches of a few microseconds each)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        y_graph = step()
    gathered = [torch.empty_like(y_graph) for _ in range(world)] if world > 1 else None

    def run_step():
        graph.replay()
        if world > 1:
            dist.all_gather(gathered, y_graph)       # the only collective: final output gather over NVLink

    for _ in range(args.warmup):
        run_step()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region 1: device-resident inputs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    w0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()         # no-op unless run under `ncu --profile-from-start off` (profiles/ launch lists)
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    torch.cuda.profiler.stop()
    barrier()
    w1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(w0, w1) if rank == 0 else None

    # ---- timed region 2: end to end with host buffers (pinned H2D of the step inputs, D2H of the step output)
    host_out = torch.empty(y_graph.shape, dtype=y_graph.dtype).pin_memory()
    for _ in range(2):
        for k_, v in host_inputs.items():
            dev_in[k_].copy_(v, non_blocking=True)
        run_step()
        host_out.copy_(y_graph, non_blocking=True)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        for k_, v in host_inputs.items():
            dev_in[k_].copy_(v, non_blocking=True)
        run_step()
        host_out.copy_(y_graph, non_blocking=True)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    h2d = sum(v.numel() * v.element_size() for v in host_inputs.values())
    d2h = host_out.numel() * host_out.element_size()

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    # ---- roofline of the dominant kernel (rank 0): every launch of that kernel in one step, replayed as a graph between CUDA events
    roofline = cpu_base = None
    if rank == 0:
        from sdnq_b200.forward import matmul_operand
        acts = derive_activations(dev_in, shapes)

        def graph_time(fn, reps=5):
            """capture fn() (a list of kernel launches) in a CUDA graph and time `reps` replays with CUDA events: per-launch
            durations without any host launch overhead in between (what the kernels cost inside the real, graph-replayed step)"""
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            g.replay()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(reps):
                g.replay()
            t1.record()
            t1.synchronize()
            return t0.elapsed_time(t1) / reps

        mm_layers = [(m, n, k, layer) for _, m, n, k, layer in stack if layer.sdnq_dequantizer.use_quantized_matmul and m >= 32]
        dq_layers = [(m, n, k, layer) for _, m, n, k, layer in stack if not layer.sdnq_dequantizer.use_quantized_matmul]
        gemm_ms = gemm_flops = k2_ms = k2_bytes = dq_ms = dq_bytes = 0.0
        n_gemm, n_dq = len(mm_layers), len(dq_layers)
        if mm_layers:
            pre = {}
            for m, n, k, layer in mm_layers:        # one pre-quantised activation per distinct (M, K, mode): L2-hot like in the step
                d = layer.sdnq_dequantizer
                key = (m, k, d.quantized_matmul_dtype, d.hadamard_group_size if d.use_hadamard else 0)
                if key not in pre:
                    pre[key] = ops.act_quant(acts[(m, k)], key[2], hadamard_group=key[3], want_rowsum=matmul_operand(layer).zp is not None)

            def all_gemms():
                for m, n, k, layer in mm_layers:
                    d = layer.sdnq_dequantizer
                    op = matmul_operand(layer)
                    xq, sx, zx, rowsum, _ = pre[(m, k, d.quantized_matmul_dtype, d.hadamard_group_size if d.use_hadamard else 0)]
                    ops.scaled_mm(xq, op.wq, sx, op.sw, layer.bias, torch.bfloat16, rowsum=rowsum, zp=op.zp, colsum=op.colsum, zx=zx)

            def all_act_quants():
                for m, n, k, layer in mm_layers:
                    d = layer.sdnq_dequantizer
                    ops.act_quant(acts[(m, k)], d.quantized_matmul_dtype, hadamard_group=d.hadamard_group_size if d.use_hadamard else 0)

            gemm_ms = graph_time(all_gemms)
            k2_ms = graph_time(all_act_quants)
            gemm_flops = sum(2.0 * m * n * k for m, n, k, _ in mm_layers)
            k2_bytes = sum(3.0 * m * k for m, n, k, _ in mm_layers)
        if dq_layers:
            def all_dequants():
                for m, n, k, layer in dq_layers:
                    layer.sdnq_dequantizer(layer.weight, layer.scale, layer.zero_point, layer.svd_up, layer.svd_down)
            dq_ms = graph_time(all_dequants)
            for m, n, k, layer in dq_layers:
                tensors = [layer.weight, layer.scale, layer.zero_point, layer.svd_up, layer.svd_down]
                dq_bytes += sum(t.numel() * t.element_size() for t in tensors if t is not None) + 2.0 * n * k
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
        bf16_sustained = float(peaks.get("bf16_tflops_sustained", bf16_peak))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        # a kernel timed inside a long, power-limited region is held against the sustained figure, a short burst against the burst one
        long_region = 5 * gemm_ms >= 200.0 or (clocks is not None and "sw_power_cap" in (clocks.get("reasons") or []))
        tensor_peak = 2.0 * (bf16_sustained if long_region else bf16_peak)
        # DRAM traffic per launch of the dominant kernel: from the committed ncu capture of this workload (profiles/traffic.json,
        # written by tools/ncu_summary.py traffic from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over one step)
        traffic = traffic_note = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
            if tr:
                traffic, traffic_note = tr.get("dominant_kernel_dram_bytes_per_launch"), tr.get("source")
        except Exception:
            pass
        if n_dq and not n_gemm:
            ach = dq_bytes / dq_ms / 1e6
            roofline = {"bound": "hbm", "kernel": "dequant_svd_kernel / dequant_kernel (K3)", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ach / hbm_peak, "traffic": traffic, "traffic_note": traffic_note, "peak_note": f"HBM copy peak, {src}", "launches": n_dq,
                        "algorithmic_bytes_per_launch": dq_bytes / n_dq,
                        "avg_launch_us": 1e3 * dq_ms / n_dq,
                        "algorithmic_bytes": "packed codes + scales (+zp) + svd factors read, bf16 weight written"}
        achieved = gemm_flops / max(gemm_ms, 1e-9) / 1e9
        if n_gemm:
            roofline = {"bound": "tensor", "kernel": "gemm_w8a8_kernel (tcgen05 kind::i8 / kind::f8f6f4, single CTA or cta_group::2 pairs)",
                        "achieved": achieved, "peak": tensor_peak,
                        "unit": "TFLOP/s", "frac": achieved / tensor_peak, "traffic": traffic, "traffic_note": traffic_note,
                        "peak_note": f"8-bit dense tensor peak taken as 2x the {src} bf16 cuBLAS figure: "
                                     f"{'sustained' if long_region else 'burst'} ({bf16_sustained if long_region else bf16_peak} TF) because the timed region is "
                                     f"{'long / power-capped' if long_region else 'short'}; burst {2 * bf16_peak:.0f} / sustained {2 * bf16_sustained:.0f} TF",
                        "algorithmic_flops_per_launch": gemm_flops / max(n_gemm, 1),
                        "algorithmic_bytes_per_launch": sum(float(n * k + m * k + 2 * m * n) for m, n, k, _ in mm_layers) / max(n_gemm, 1),
                        "launches": n_gemm, "avg_launch_us": 1e3 * gemm_ms / max(n_gemm, 1), "share_of_step": gemm_ms / (gemm_ms + k2_ms),
                        "act_quant": {"bound": "hbm", "achieved": k2_bytes / k2_ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": k2_bytes / k2_ms / 1e6 / hbm_peak, "avg_launch_us": 1e3 * k2_ms / max(n_gemm, 1)}}
        if world == 1 and not args.no_cpu_baseline:
            cpu_base = cpu_oracle_sample(args.workload, budget_s=12.0)

    if rank == 0:
        tfl = flops_step * args.steps * world / (ms * 1e-3) / 1e12
        tfl_e2e = flops_step * args.steps * world / (ms_e2e * 1e-3) / 1e12
        line = {"metric": "quantized_linear_tflops", "value": tfl, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "steps_per_s": 1e3 * args.steps / ms * world, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": spec["dtype"], "data": "synthetic",
                "config": {"workload": args.workload, "describe": spec["describe"], "linears_per_step": len(stack),
                           "gflop_per_step": flops_step / 1e9, "batch_per_gpu": 1 if args.workload == "sdxl_int8" else 4, "parallelism": f"dp{world}",
                           "l2": "each Linear has its own weights (2.2 GB int8 for sdxl_int8, 12 GB fp8 for flux_fp8) >> 126 MB L2: weights stream from HBM every step, no flush needed",
                           "execution": "whole step captured in one CUDA graph, replayed per step"},
                "e2e": {"value": tfl_e2e, "unit": "TFLOP/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches_per_step) * args.steps, "launches_per_step": int(launches_per_step),
                "clocks": clocks, "roofline": roofline}
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="sdxl_int8", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
