"""Benchmark of the quantized-Linear hot path (contract: see the repo brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One "step" = one pass over every SDNQ-quantised Linear of one denoise step of the named model on synthetic activations of
the real shapes, each layer with its own random weights quantised through the public API
(`sdnq_quantize_layer(nn.Linear, SDNQConfig)`), run through `SDNQLinear.forward` -> C ABI -> sm_100a kernels.  Layers that share
an input in the real model (to_q / to_k / to_v of one attention block, the cross-attention to_k / to_v of every block, FLUX
single-block proj_mlp) are fed the same tensor; every other layer gets its own buffer.

Headline (default) workload: `flux_fp8` = BASELINE.json configs[2] (FLUX.1-dev fp8 + Hadamard(256), 4 images per GPU; with
--gpus N it is configs[4]: batch 4N sharded over N GPUs).  Without --workload the same run also measures `sdxl_int8`
(configs[1]) and `sdxl_int4_svd_dequant` (configs[3]) and nests their lines under "workloads".

  value         whole-job TFLOP/s (2*M*N*K over all Linears of the step), inputs resident in HBM, step replayed as a CUDA graph
  e2e           same metric through the same call with HOST inputs: pinned host -> device copy of the step inputs and a device ->
                host read of the step output inside the timed region
  peaks         8-bit dense tensor peak measured in this run (cuBLASLt int8 / fp8 8192^3, burst and sustained)
  roofline      the dominant kernel: algorithmic FLOPs (bytes) of all its launches in a step / their CUDA-event time
                (the launches replayed back to back as a graph), against the measured peak
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, a scripted copy of its package) on the host cores: CPU-eager path,
                all cores, bounded sample of the same workload
  gpu_reference the reference's own GPU paths on this B200 (Triton scaled-mm kernel; CUDA eager = cuBLASLt), same layer shapes

N > 1 (torchrun): every rank runs the same stack on its own batch shard (weights replicated, no data-path collective), one
NCCL all_gather of the step output per step; time = max over ranks; scaling = weak.
`--impl reference`: the reference's CPU-eager path on the host cores (rank 0 only), same metric / config.
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
# a layer is (name, M, N, K, src): src names the tensor that feeds it -- layers with the same src share one input buffer
def sdxl_linears(batch=1):
    """SD-XL UNet, 1024x1024 (latents 128^2), `batch` images: 4096 tokens per image at C = 640 (10 transformer layers in 5
    Transformer2DModels), 1024 at C = 1280 (60 layers in 6 models), 77 text tokens of width 2048 (one encoder_hidden_states
    tensor feeds the cross-attention to_k / to_v of every block); plus the M = batch Linears the reference also quantises
    (ResNet time_emb_proj, add_embedding), which take the small-M path."""
    layers = []
    for C, tokens, models, per_model in ((640, 4096, 5, 2), (1280, 1024, 6, 10)):
        M = tokens * batch
        for t in range(models):
            layers.append((f"c{C}.t{t}.proj_in", M, C, C, f"c{C}.t{t}.in"))
            for blk in range(per_model):
                p = f"c{C}.t{t}.b{blk}"
                layers += [(f"{p}.attn1.to_q", M, C, C, f"{p}.norm1"), (f"{p}.attn1.to_k", M, C, C, f"{p}.norm1"), (f"{p}.attn1.to_v", M, C, C, f"{p}.norm1"),
                           (f"{p}.attn1.to_out", M, C, C, f"{p}.attn1.o"), (f"{p}.attn2.to_q", M, C, C, f"{p}.norm2"),
                           (f"{p}.attn2.to_k", 77 * batch, C, 2048, "context"), (f"{p}.attn2.to_v", 77 * batch, C, 2048, "context"),
                           (f"{p}.attn2.to_out", M, C, C, f"{p}.attn2.o"), (f"{p}.ff.proj", M, 8 * C, C, f"{p}.norm3"), (f"{p}.ff.out", M, C, 4 * C, f"{p}.ff.h")]
            layers.append((f"c{C}.t{t}.proj_out", M, C, C, f"c{C}.t{t}.out"))
    for i, c_out in enumerate([320] * 2 + [640] * 2 + [1280] * 2 + [1280] * 2 + [1280] * 3 + [640] * 3 + [320] * 3):
        layers.append((f"resnet{i}.time_emb_proj", batch, c_out, 1280, f"resnet{i}.temb"))
    layers += [("add_embedding.linear_1", batch, 1280, 2816, "add_emb.in"), ("add_embedding.linear_2", batch, 1280, 1280, "add_emb.h")]
    return layers


def flux_linears(batch=4):
    """FLUX.1-dev DiT, 1024x1024, `batch` images: 4096 image tokens + 512 text tokens per image, D = 3072; 19 double blocks
    (separate image / text streams; q / k / v of a stream share the normed hidden states), 38 single blocks (q / k / v / proj_mlp
    share them); AdaLN Linears have M = batch (small-M path)."""
    D, layers = 3072, []
    mi, mt = 4096 * batch, 512 * batch
    for b in range(19):
        for stream, M in (("img", mi), ("txt", mt)):
            p = f"double{b}.{stream}"
            layers += [(f"{p}.norm1.linear", batch, 6 * D, D, f"double{b}.temb"), (f"{p}.to_q", M, D, D, f"{p}.norm1"), (f"{p}.to_k", M, D, D, f"{p}.norm1"),
                       (f"{p}.to_v", M, D, D, f"{p}.norm1"), (f"{p}.to_out", M, D, D, f"{p}.attn.o"), (f"{p}.ff.proj", M, 4 * D, D, f"{p}.norm2"),
                       (f"{p}.ff.out", M, D, 4 * D, f"{p}.ff.h")]
    for b in range(38):
        p, M = f"single{b}", mi + mt
        if b > 0:
            layers.append((f"{p}.norm.linear", batch, 3 * D, D, f"{p}.temb"))
        layers += [(f"{p}.to_q", M, D, D, f"{p}.norm"), (f"{p}.to_k", M, D, D, f"{p}.norm"), (f"{p}.to_v", M, D, D, f"{p}.norm"),
                   (f"{p}.proj_mlp", M, 4 * D, D, f"{p}.norm"), (f"{p}.proj_out", M, D, 5 * D, f"{p}.cat")]
    return layers


WORKLOADS = {
    "flux_fp8": dict(describe="FLUX.1-dev DiT float8_e4m3fn + Hadamard(256) W8A8, 4 images per GPU, 1024x1024: all quantised Linears of one denoise step "
                              "(BASELINE configs[2]; with N GPUs configs[4]: batch 4N sharded)",
                     layers=flux_linears, batch_per_gpu=4, dtype="f8e4m3", model="FLUX.1-dev",
                     config=dict(weights_dtype="float8_e4m3fn", use_quantized_matmul=True, use_hadamard=True, hadamard_group_size=256)),
    "sdxl_int8": dict(describe="SD-XL UNet int8 W8A8 (use_quantized_matmul=True), 1 image per GPU, 1024x1024: all quantised Linears of one denoise step "
                               "(BASELINE configs[1])",
                      layers=sdxl_linears, batch_per_gpu=1, dtype="int8", model="SD-XL UNet",
                      config=dict(weights_dtype="int8", use_quantized_matmul=True)),
    "sdxl_int4_svd_dequant": dict(describe="SD-XL UNet int4 group_size=128 + SVD rank 32, dequant path (use_quantized_matmul=False), 1 image per GPU, 1024x1024 "
                                           "(BASELINE configs[3])",
                                  layers=sdxl_linears, batch_per_gpu=1, dtype="bf16", model="SD-XL UNet",
                                  config=dict(weights_dtype="int4", group_size=128, use_svd=True, svd_rank=32)),
}
HEADLINE = "flux_fp8"
NESTED = ("sdxl_int8", "sdxl_int4_svd_dequant")


def total_flops(layers):
    return sum(2.0 * m * n * k for _, m, n, k, _ in layers)


def workload_config(name, world):
    """The workload definition, identical for every arm (ours / reference): nothing device- or implementation-specific in it."""
    spec = WORKLOADS[name]
    layers = spec["layers"](spec["batch_per_gpu"])
    weight_gb = sum(n * k for _, _, n, k, _ in layers) * (0.5 if "int4" in name else 1.0) / 1e9
    return {"workload": name, "describe": spec["describe"], "model_shapes": spec["model"], "sdnq_config": spec["config"],
            "linears_per_step": len(layers), "gflop_per_step_per_gpu": total_flops(layers) / 1e9,
            "batch_per_gpu": spec["batch_per_gpu"], "global_batch": spec["batch_per_gpu"] * world, "parallelism": f"dp{world}",
            "l2": f"each Linear has its own weights ({weight_gb:.1f} GB per step >> 126 MB L2) and non-sibling layers their own activation buffers: "
                  "weights stream from HBM every step, no flush needed"}


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


# ------------------------------------------------------------------------------------------------ the reference itself
CPU_M_CAP = 4096


def reference_sample_shapes(workload, m_cap=None):
    """One layer per distinct (M, N, K) of the workload's W8A8 / dequant-path GEMMs (M >= 32), with its multiplicity."""
    spec = WORKLOADS[workload]
    count = {}
    for _, m, n, k, _ in spec["layers"](spec["batch_per_gpu"]):
        if m >= 32:
            key = (min(m, m_cap) if m_cap else m, n, k)
            count[key] = count.get(key, 0) + 1
    return sorted(count.items())


def _reference_layers(workload, device, dtype, m_cap):
    """Quantise one nn.Linear per sample shape with the REFERENCE's own sdnq_quantize_layer and SDNQConfig."""
    import torch

    from oracle.ref_loader import load_reference
    sdnq = load_reference()
    from sdnq.quantizer import sdnq_quantize_layer as ref_quantize_layer
    spec = WORKLOADS[workload]
    torch.manual_seed(1234)
    out = []
    for (m, n, k), mult in reference_sample_shapes(workload, m_cap):
        lin = torch.nn.Linear(k, n, bias=True, device=device, dtype=dtype)
        layer = ref_quantize_layer(lin, sdnq.SDNQConfig(**spec["config"]))[0]
        x = torch.randn(m, k, device=device, dtype=dtype)
        out.append((m, n, k, mult, layer, x))
    return out


def cpu_reference_run(workload, steps, warmup, budget_s=None):
    """The unmodified reference on the host cores: CPU-eager path (SDNQ_DEVICE=cpu, SDNQ_USE_TORCH_COMPILE=0; model dtype fp32
    as sdnext.py:44 picks for CPU), every host thread, one step = one forward of one layer per distinct shape of the workload with
    M capped at CPU_M_CAP rows.  -> (TFLOP/s, seconds per step, cpu_baseline dict)."""
    os.environ["SDNQ_DEVICE"] = "cpu"
    os.environ["SDNQ_USE_TORCH_COMPILE"] = "0"
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)                    # torchrun exports OMP_NUM_THREADS=1: override it for the host arm
    layers = _reference_layers(workload, torch.device("cpu"), torch.float32, CPU_M_CAP)
    flops = sum(2.0 * m * n * k for m, n, k, _, _, _ in layers)

    def one_step():
        with torch.no_grad():
            for _, _, _, _, layer, x in layers:
                layer(x)
    t_w0 = time.perf_counter()
    for _ in range(max(warmup, 1)):
        one_step()
    per_step = (time.perf_counter() - t_w0) / max(warmup, 1)
    if budget_s is not None:
        steps = max(1, min(steps, int(budget_s / max(per_step, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    value = flops * steps / dt / 1e12
    from oracle.ref_loader import reference_root
    info = {"value": value, "unit": "TFLOP/s", "cores": cores, "threads": torch.get_num_threads(), "kind": "reference",
            "sample": f"{len(layers)} layers (one per distinct MxNxK of {workload}, M capped at {CPU_M_CAP} rows) x {steps} timed passes after "
                      f"{max(warmup, 1)} warm-up, unmodified reference ({os.path.relpath(reference_root() or '?', ROOT)}) CPU-eager path, fp32 activations, "
                      f"{dt:.1f} s; FLOPs counted = those computed"}
    return value, dt / steps, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = args.workload or HEADLINE
    spec = WORKLOADS[name]
    value, s_per_step, info = cpu_reference_run(name, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "quantized_linear_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": spec["dtype"], "data": "synthetic", "config": workload_config(name, world),
            "cpu_baseline": info, "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_gpu_reference_arm(args):
    """Child process of the GPU arm: the reference's own GPU path on cuda:0 (flags come from the environment the parent set),
    one layer per distinct shape, CUDA events around 10 back-to-back forwards after 3 warm-ups; the step time is extrapolated with
    the shape multiplicities.  Prints one JSON object per workload."""
    import torch

    from oracle.ref_loader import load_reference
    load_reference()
    dev = torch.device("cuda", 0)
    try:
        # the reference's Triton kernels build their TMA descriptors on the device (tl.make_tensor_descriptor), which needs a
        # scratch allocator from the host program; Inductor installs one for compiled graphs, a plain eager caller has to
        import triton
        triton.set_allocator(lambda size, align, stream: torch.empty(size, dtype=torch.int8, device=dev))
    except Exception:      # noqa: BLE001
        pass
    out = {}
    for name in args.workload.split(","):
        layers = _reference_layers(name, dev, torch.bfloat16, None)
        rows, step_ms, flops = [], 0.0, 0.0
        with torch.no_grad():
            for m, n, k, mult, layer, x in layers:
                for _ in range(3):
                    layer(x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    layer(x)
                e1.record()
                e1.synchronize()
                ms = e0.elapsed_time(e1) / 10
                rows.append({"M": m, "N": n, "K": k, "count": mult, "ms": ms, "tflops": 2.0 * m * n * k / ms / 1e9})
                step_ms += mult * ms
                flops += mult * 2.0 * m * n * k
        out[name] = {"value": flops / step_ms / 1e9, "unit": "TFLOP/s", "step_ms_extrapolated": step_ms, "per_shape": rows,
                     "forward_func": layers[0][4].forward_func.__name__}
    print("GPU_REFERENCE " + json.dumps(out), flush=True)


GPU_REFERENCE_VARIANTS = {
    # the reference's defaults on a B200 with torch.compile off: Triton scaled-mm kernel (kernels/triton_scaled_mm.py:239-275)
    "triton": {"SDNQ_USE_TORCH_COMPILE": "0"},
    # CUDA eager: torch._int_mm / torch._scaled_mm (cuBLASLt) + elementwise epilogue (kernel_wrappers.py:132-150)
    "cuda_eager": {"SDNQ_USE_TORCH_COMPILE": "0", "SDNQ_USE_TRITON_MM": "0"},
}


def gpu_reference(workloads, timeout_s=420):
    """Run the reference's GPU paths in child processes (its flags are resolved at import time) and collect their numbers."""
    from oracle.ref_loader import reference_root
    if reference_root() is None:
        return {"unavailable": "oracle/_ref is missing (python oracle/build_ref.py in the authoring container)"}
    res = {"how": "unmodified reference on cuda:0, bf16, one layer per distinct MxNxK, CUDA events around 10 forwards after 3 warm-ups (eager launches, "
                  "no CUDA graph: the reference has none), step time extrapolated with the shape multiplicities; the harness installs a "
                  "triton.set_allocator scratch allocator (the reference's kernels build TMA descriptors on the device)"}
    for variant, env in GPU_REFERENCE_VARIANTS.items():
        e = dict(os.environ, SDNQ_DEVICE="cuda", CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0], **env)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            e.pop(k, None)
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "gpu_reference", "--workload", ",".join(workloads)],
                               capture_output=True, text=True, timeout=timeout_s, env=e, cwd=ROOT)
            got = [ln for ln in p.stdout.splitlines() if ln.startswith("GPU_REFERENCE ")]
            if got:
                res[variant] = json.loads(got[-1][len("GPU_REFERENCE "):])
                res[variant]["seconds"] = round(time.time() - t0, 1)
            else:
                res[variant] = {"unavailable": (p.stderr or p.stdout).strip().splitlines()[-1][:300] if (p.stderr or p.stdout).strip() else f"exit {p.returncode}"}
        except subprocess.TimeoutExpired:
            res[variant] = {"unavailable": f"timed out after {timeout_s} s"}
        except Exception as ex:      # noqa: BLE001
            res[variant] = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}
    return res


# ------------------------------------------------------------------------------------------------ measured tensor peaks
def measure_peaks(device):
    """Dense 8-bit tensor-core peak of this GPU, measured the way MEASURED_PEAKS.json measures bf16: the library GEMM at 8192^3
    (torch._int_mm = cuBLASLt int8 -> int32; torch._scaled_mm = cuBLASLt e4m3 x e4m3 -> bf16), best of 10 single launches
    (burst) and back to back for ~1.5 s (sustained)."""
    import torch
    n = 8192
    flops = 2.0 * n ** 3
    out = {"how": "torch._int_mm (int8 -> int32) and torch._scaled_mm (e4m3 -> bf16, tensor-wise unit scales) at 8192^3 on this GPU in this run: "
                  "best of 10 launches (burst), back to back for 1.5 s (sustained)"}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(1500.0 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        sustained = flops * reps / e0.elapsed_time(e1) / 1e9
        return max(flops / best / 1e9, sustained), sustained        # (a cold single launch can time below the warmed-up loop)
    try:
        a = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=device)
        b = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=device).t()
        out["int8_tflops"], out["int8_tflops_sustained"] = timed(lambda: torch._int_mm(a, b))
    except Exception as ex:      # noqa: BLE001
        out["int8_error"] = f"{type(ex).__name__}: {ex}"[:200]
    try:
        a8 = torch.randn(n, n, device=device).to(torch.float8_e4m3fn)
        b8 = torch.randn(n, n, device=device).to(torch.float8_e4m3fn).t()
        one = torch.ones((), device=device, dtype=torch.float32)
        out["fp8_tflops"], out["fp8_tflops_sustained"] = timed(lambda: torch._scaled_mm(a8, b8, scale_a=one, scale_b=one, out_dtype=torch.bfloat16))
    except Exception as ex:      # noqa: BLE001
        out["fp8_error"] = f"{type(ex).__name__}: {ex}"[:200]
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def build_stack(workload, device):
    import torch

    from sdnq_b200 import SDNQConfig, sdnq_quantize_layer
    spec = WORKLOADS[workload]
    cfg = spec["config"]
    stack = []
    torch.manual_seed(1234)
    for name, m, n, k, src in spec["layers"](spec["batch_per_gpu"]):
        lin = torch.nn.Linear(k, n, bias=True, device=device, dtype=torch.bfloat16)
        layer, _ = sdnq_quantize_layer(lin, SDNQConfig(**cfg), param_name=name + ".weight")
        stack.append((name, m, n, k, src, layer))
    return stack


POOL = 3      # distinct buffers per activation shape, handed out round-robin to the distinct producers


def make_activations(stack, dev_in, device):
    """src -> activation tensor.  The step inputs (`dev_in`: "hidden" = the input of the first token-level Linear, "context" = the
    text-encoder states of SD-XL) feed their consumers directly; every other producer (attention output, norm output, MLP hidden
    state ...) is a stand-in buffer of the right shape, distinct from the buffers of the producers around it, so that only layers
    that really share an input in the model present the same tensor to the library."""
    import torch
    acts, pools, nxt = {}, {}, {}
    first = next(src for _, m, _, _, src, _ in stack if m >= 32)
    for _, m, _, k, src, _ in stack:
        if src in acts:
            continue
        if src == first:
            acts[src] = dev_in["hidden"]
        elif src == "context" and "context" in dev_in:
            acts[src] = dev_in["context"]
        else:
            pool = pools.setdefault((m, k), [])
            i = nxt.get((m, k), 0)
            if len(pool) < POOL:
                pool.append(torch.randn(m, k, device=device, dtype=torch.bfloat16))
            acts[src] = pool[i % POOL]
            nxt[(m, k)] = i + 1
    return acts


def run_workload(name, args, world, rank, device, peaks, with_roofline=True):
    """Build, warm up, capture and time one workload.  Returns the result dict (rank 0) or None."""
    import torch
    import torch.distributed as dist

    from sdnq_b200 import _lib, ops
    spec = WORKLOADS[name]
    local_rank = device.index
    stack = build_stack(name, device)
    sib_groups = 0
    if args.siblings != "off":      # what sdnq_post_load_quant does on a module tree, here on the (name, layer) list of the synthetic stack
        from sdnq_b200 import fuse_named_siblings
        sib_groups = fuse_named_siblings([(n, layer) for n, _, _, _, _, layer in stack], cross_attention_pool=8 if args.siblings == "pool" else 0)
    layers_meta = [(n, m, nn_, k, s) for n, m, nn_, k, s, _ in stack]
    flops_step = total_flops(layers_meta)

    # ---- step inputs live on the host (pinned); everything else is derived on device
    torch.manual_seed(100 + rank)
    hidden_shape = next((m, k) for _, m, _, k, _ in layers_meta if m >= 32)      # the first Linear's input: the step's "latents"
    host_inputs = {"hidden": torch.randn(hidden_shape, dtype=torch.bfloat16).pin_memory()}
    ctx = [(m, k) for _, m, _, k, s in layers_meta if s == "context"]
    if ctx:
        host_inputs["context"] = torch.randn(ctx[0], dtype=torch.bfloat16).pin_memory()
    dev_in = {k: v.to(device, non_blocking=True) for k, v in host_inputs.items()}
    acts = make_activations(stack, dev_in, device)
    out_holder = {}

    def step():
        for _, m, _, _, src, layer in stack:
            y = layer(acts[src])
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

    # ---- capture the whole step in one CUDA graph (launch-bound inner loop)
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

    # ---- timed region 2: end to end with host buffers.  Every step copies its inputs from pinned host memory and its output back to
    # pinned host memory, all inside the timed region; as a serving loop would, the copies run on two copy streams through device staging
    # buffers, so the H2D of step i+1 and the D2H of step i-1 overlap the compute of step i (the graph reads / writes its static buffers
    # through a device-to-device copy on the compute stream).  The first H2D and the last D2H are not hidden and are inside the region.
    host_out = torch.empty(y_graph.shape, dtype=y_graph.dtype).pin_memory()
    main_s = torch.cuda.current_stream()
    h2d_s, d2h_s = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
    stage_in = {k_: torch.empty_like(v) for k_, v in dev_in.items()}
    stage_out = torch.empty_like(y_graph)

    def e2e_steps(n):
        def upload(after=None):
            with torch.cuda.stream(h2d_s):
                if after is not None:
                    h2d_s.wait_event(after)                    # the staging buffers have been read by the previous step
                for k_, v in host_inputs.items():
                    stage_in[k_].copy_(v, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(h2d_s)
            return ev
        h2d_s.wait_stream(main_s)
        d2h_s.wait_stream(main_s)
        landed, drained = upload(), None
        for i in range(n):
            main_s.wait_event(landed)                          # this step's inputs are in the staging buffers
            for k_ in dev_in:
                dev_in[k_].copy_(stage_in[k_], non_blocking=True)
            taken = torch.cuda.Event()
            taken.record(main_s)
            if i + 1 < n:
                landed = upload(after=taken)                   # next step's inputs travel while this step computes
            run_step()
            if drained is not None:
                main_s.wait_event(drained)                     # the previous result has left the output staging buffer
            stage_out.copy_(y_graph, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(main_s)
            with torch.cuda.stream(d2h_s):
                d2h_s.wait_event(ready)
                host_out.copy_(stage_out, non_blocking=True)   # this step's result travels while the next step computes
                drained = torch.cuda.Event()
                drained.record(d2h_s)
        main_s.wait_event(drained)                             # the last result is on the host before the region ends

    e2e_steps(2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_steps(args.steps)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    h2d = sum(v.numel() * v.element_size() for v in host_inputs.values())
    d2h = host_out.numel() * host_out.element_size()

    if world > 1:      # a multi-GPU step takes as long as its slowest rank
        from sdnq_b200.parallel import max_over_ranks
        ms, ms_e2e = max_over_ranks(ms, device=device), max_over_ranks(ms_e2e, device=device)

    # ---- roofline of the dominant kernel (rank 0): every launch of that kernel in one step, replayed as a graph between CUDA events
    roofline = None
    if rank == 0 and with_roofline:
        from sdnq_b200.forward import matmul_operand

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

        mm_layers = [(m, n, k, src, layer) for _, m, n, k, src, layer in stack if layer.sdnq_dequantizer.use_quantized_matmul and m >= 32]
        dq_layers = [(m, n, k, src, layer) for _, m, n, k, src, layer in stack if not layer.sdnq_dequantizer.use_quantized_matmul and m >= 32]
        gemm_ms = gemm_flops = k2_ms = k2_bytes = dq_ms = dq_bytes = 0.0
        n_gemm, n_dq, n_k2 = len(mm_layers), len(dq_layers), 0
        peaks_file = {}
        try:
            peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:      # noqa: BLE001
            pass
        hbm_peak = float(peaks_file.get("hbm_gbs", 6650.0))
        src_note = "measured (MEASURED_PEAKS.json)" if peaks_file else "fallback (B200_PROFILING.md)"
        traffic = traffic_note = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
            if tr:
                traffic, traffic_note = tr.get("dominant_kernel_dram_bytes_per_launch"), tr.get("source")
        except Exception:      # noqa: BLE001
            pass
        if mm_layers:
            pre, k2_srcs, last = {}, [], None
            for m, n, k, src, layer in mm_layers:        # one pre-quantised activation per distinct source tensor
                d = layer.sdnq_dequantizer
                key = (src, d.quantized_matmul_dtype, d.hadamard_group_size if d.use_hadamard else 0)
                if key not in pre:
                    pre[key] = ops.act_quant(acts[src], key[1], hadamard_group=key[2], want_rowsum=matmul_operand(layer).zp is not None)
                if key != last:                          # the K2 launches the step really makes: one per run of siblings
                    k2_srcs.append((m, k, key))
                    last = key

            # the GEMM launches the step really makes: one per layer, or one per sibling group and input tensor
            gemm_calls, seen = [], set()
            for m, n, k, src, layer in mm_layers:
                d = layer.sdnq_dequantizer
                key = (src, d.quantized_matmul_dtype, d.hadamard_group_size if d.use_hadamard else 0)
                group = layer.__dict__.get("_sdnq_siblings")
                if group is not None and group.state is not None and not group.dead:
                    if (id(group), src) not in seen:
                        seen.add((id(group), src))
                        gemm_calls.append((key, None, group.state))
                else:
                    gemm_calls.append((key, layer, None))

            def all_gemms():
                for key, layer, st in gemm_calls:
                    xq, sx, zx, rowsum, _ = pre[key]
                    if st is not None:
                        ops.scaled_mm_grouped(xq, st["wq"], sx, st["sw"], st["starts"], st["ns"], st["bias"], torch.bfloat16, rowsum=rowsum,
                                              zp=st["zp"], colsum=st["colsum"], zx=zx, packed_dtype=st["packed"])
                    else:
                        op = matmul_operand(layer)
                        ops.scaled_mm(xq, op.wq, sx, op.sw, layer.bias, torch.bfloat16, rowsum=rowsum, zp=op.zp, colsum=op.colsum, zx=zx)

            def all_act_quants():
                for m, k, (src, mmd, hg) in k2_srcs:
                    ops.act_quant(acts[src], mmd, hadamard_group=hg)

            gemm_ms = graph_time(all_gemms)
            k2_ms = graph_time(all_act_quants)
            n_k2 = len(k2_srcs)
            n_gemm = len(gemm_calls)
            gemm_flops = sum(2.0 * m * n * k for m, n, k, _, _ in mm_layers)
            k2_bytes = sum(3.0 * m * k for m, k, _ in k2_srcs)
            fp8 = spec["config"]["weights_dtype"].startswith("float")
            kind = "fp8" if fp8 else "int8"
            burst, sustained = peaks.get(f"{kind}_tflops"), peaks.get(f"{kind}_tflops_sustained")
            if burst is None:       # library peak not measurable: 2 x the bf16 figures
                burst = 2.0 * float(peaks_file.get("bf16_tflops", 1590.0))
                sustained = 2.0 * float(peaks_file.get("bf16_tflops_sustained", burst / 2.0))
                peak_src = f"2x the bf16 cuBLAS figure, {src_note}"
            else:
                peak_src = f"cuBLASLt {kind} 8192^3 measured in this run"
            # a kernel timed inside a long, power-limited region is held against the sustained figure, a short burst against the burst one
            long_region = 5 * gemm_ms >= 200.0 or (clocks is not None and "sw_power_cap" in (clocks.get("reasons") or []))
            tensor_peak = sustained if long_region else burst
            achieved = gemm_flops / max(gemm_ms, 1e-9) / 1e9
            roofline = {"bound": "tensor", "kernel": "gemm_w8a8_kernel (tcgen05 kind::i8 / kind::f8f6f4, single CTA or cta_group::2 pairs)",
                        "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                        "traffic": traffic, "traffic_note": traffic_note,
                        "peak_note": f"{peak_src}: {'sustained' if long_region else 'burst'} figure because the timed region is "
                                     f"{'long / power-capped' if long_region else 'short'} (burst {burst:.0f} / sustained {sustained:.0f} TF)",
                        "algorithmic_flops_per_launch": gemm_flops / max(n_gemm, 1),
                        "algorithmic_bytes_per_launch": sum(float(n * k + m * k + 2 * m * n) for m, n, k, _, _ in mm_layers) / max(n_gemm, 1),
                        "launches": n_gemm, "avg_launch_us": 1e3 * gemm_ms / max(n_gemm, 1), "share_of_step": gemm_ms / (gemm_ms + k2_ms),
                        "act_quant": {"bound": "hbm", "kernel": "act_quant_kernel (K2)", "achieved": k2_bytes / k2_ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": k2_bytes / k2_ms / 1e6 / hbm_peak, "launches": n_k2, "avg_launch_us": 1e3 * k2_ms / max(n_k2, 1),
                                      "launches_without_sibling_reuse": len(mm_layers)},
                        "linears": len(mm_layers), "grouped_launches": sum(1 for c in gemm_calls if c[2] is not None)}
        elif dq_layers:
            roofline = dequant_path_roofline(dq_layers, acts, graph_time, hbm_peak, src_note, peaks_file, traffic, traffic_note)

    if rank != 0:
        return None
    tfl = flops_step * args.steps * world / (ms * 1e-3) / 1e12
    tfl_e2e = flops_step * args.steps * world / (ms_e2e * 1e-3) / 1e12
    res = {"metric": "quantized_linear_tflops", "value": tfl, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "steps_per_s": 1e3 * args.steps / ms * world, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": spec["dtype"], "data": "synthetic", "config": workload_config(name, world),
           "execution": "whole step captured in one CUDA graph, replayed per step; sibling projections (to_q / to_k / to_v, cross-attention to_k / to_v) "
                        f"share one activation-quantise launch and one grouped GEMM launch (--siblings {args.siblings}: {sib_groups} groups)",
           "e2e": {"value": tfl_e2e, "unit": "TFLOP/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "copies": "pinned host <-> device staging buffers on two copy streams, overlapped with the neighbouring steps' compute; "
                             "first upload and last download inside the timed region"},
           "gpu_launches": int(launches_per_step) * args.steps, "launches_per_step": int(launches_per_step),
           "clocks": clocks, "roofline": roofline}
    del graph, stack, acts
    torch.cuda.empty_cache()
    return res


def dequant_path_roofline(dq_layers, acts, graph_time, hbm_peak, src_note, peaks_file, traffic, traffic_note):
    """Dequant-path workload: the dominant kernel is whatever the forward launches per layer.  With the fused W4A16 kernel the
    weight never reaches HBM as bf16, so the kernel is a GEMM (tensor bound, bf16 rate); with dequantise + library GEMM it is the
    dequant kernel (HBM bound).  Both are timed; the one the forward uses is the headline of this line."""
    import torch
    n_dq = len(dq_layers)

    def all_forwards():
        for m, n, k, src, layer in dq_layers:
            layer(acts[src])

    def all_dequants():
        for m, n, k, src, layer in dq_layers:
            layer.sdnq_dequantizer(layer.weight, layer.scale, layer.zero_point, layer.svd_up, layer.svd_down)
    fwd_ms = graph_time(all_forwards)
    dq_ms = graph_time(all_dequants)
    dq_bytes = 0.0
    for m, n, k, src, layer in dq_layers:
        tensors = [layer.weight, layer.scale, layer.zero_point, layer.svd_up, layer.svd_down]
        dq_bytes += sum(t.numel() * t.element_size() for t in tensors if t is not None) + 2.0 * n * k
    flops = sum(2.0 * m * n * k for m, n, k, _, _ in dq_layers)
    bf16_peak = float(peaks_file.get("bf16_tflops", 1590.0))
    ach_dq = dq_bytes / dq_ms / 1e6
    k3 = {"bound": "hbm", "kernel": "dequant_svd_kernel / dequant_kernel (K3: the stand-alone dequantiser, used by dequantize() and as the fallback)",
          "achieved": ach_dq, "peak": hbm_peak, "unit": "GB/s", "frac": ach_dq / hbm_peak, "launches": n_dq,
          "algorithmic_bytes_per_launch": dq_bytes / n_dq, "avg_launch_us": 1e3 * dq_ms / n_dq, "peak_note": f"HBM copy peak, {src_note}",
          "algorithmic_bytes": "packed codes + scales (+zp) + svd factors read, bf16 weight written"}
    from sdnq_b200 import forward as F
    from sdnq_b200 import ops, prefetch
    fused = bool(getattr(F, "w4a16_enabled", lambda: False)())
    dev = dq_layers[0][4].weight.device
    pf = prefetch._PREFETCHERS.get(dev.index) if prefetch.enabled() else None
    if not fused and pf is not None and pf.plans:
        # what the step really launches: the learned chains of layers, each dequantised by one batched launch (prefetch.py)
        entries = [e for e in pf.plans.values() if e.good]
        covered = {id(st.refs[0]()) for e in entries for st in e.states if st.refs[0]() is not None}

        def all_batches():
            for e in entries:
                ops.dequant_batch_run(e.plans[0])
        b_ms = graph_time(all_batches)
        b_bytes = 0.0
        for m, n, k, src, layer in dq_layers:
            if id(layer.weight) in covered:
                tensors = [layer.weight, layer.scale, layer.zero_point, layer.svd_up, layer.svd_down]
                b_bytes += sum(t.numel() * t.element_size() for t in tensors if t is not None) + 2.0 * n * k
        ach_b = b_bytes / b_ms / 1e6
        return {"bound": "hbm", "kernel": "dequant_svd_kernel, batched (K3s over the weights of the next layers in the learned call order: one persistent launch "
                                          "per chain, on the side stream ahead of the GEMMs)",
                "achieved": ach_b, "peak": hbm_peak, "unit": "GB/s", "frac": ach_b / hbm_peak, "launches": len(entries),
                "layers_covered": len(covered), "layers": n_dq, "algorithmic_bytes_per_launch": b_bytes / max(len(entries), 1),
                "avg_launch_us": 1e3 * b_ms / max(len(entries), 1), "peak_note": f"HBM copy peak, {src_note}",
                "algorithmic_bytes": "packed codes + scales (+zp) + svd factors read, bf16 weight written",
                "traffic": traffic, "traffic_note": traffic_note, "forward_ms_all_layers": fwd_ms, "per_layer_launches": k3}
    if not fused:
        return dict(k3, traffic=traffic, traffic_note=traffic_note, forward_ms_all_layers=fwd_ms)
    ach = flops / fwd_ms / 1e9
    return {"bound": "tensor", "kernel": "gemm_w4a16_kernel (tcgen05 kind::f16, packed int4 dequantised in the GEMM prologue, SVD rank-r second accumulate)",
            "achieved": ach, "peak": bf16_peak, "unit": "TFLOP/s", "frac": ach / bf16_peak, "traffic": traffic, "traffic_note": traffic_note,
            "peak_note": f"bf16 cuBLAS burst figure, {src_note} (the activations are bf16: the contraction runs at the 16-bit tensor rate)",
            "algorithmic_flops_per_launch": flops / n_dq, "launches": n_dq, "avg_launch_us": 1e3 * fwd_ms / n_dq,
            "algorithmic_bytes_per_launch": sum(float(sum(t.numel() * t.element_size() for t in (L.weight, L.scale, L.svd_up, L.svd_down) if t is not None)
                                                      + 2 * m * k + 2 * m * n) for m, n, k, _, L in dq_layers) / n_dq,
            "dequant_kernel": k3}


# ------------------------------------------------------------------------------------------------ quantized attention
ATTENTION = dict(name="flux_attention_int8", calls=57, heads=24, tokens=4608, head_dim=128, batch_per_gpu=4,
                 describe="FLUX.1-dev joint attention with SDNQ quantized Q.K^T (int8 codes, smooth-K, 16-bit P.V = sdnq_triton_atten defaults): the "
                          "57 attention calls (19 double + 38 single blocks) of one denoise step, 4 images per GPU, 24 heads x 4608 tokens x 128")


def attention_reference_child():
    """`--impl attention_reference`: the unmodified reference's Triton attention (oracle/_ref, full autotune space) on the same shape."""
    import torch
    sys.path.insert(0, ROOT)
    a = ATTENTION
    from oracle.ref_loader import load_reference
    load_reference(SDNQ_DEVICE="cuda", SDNQ_USE_TORCH_COMPILE="0")
    import triton
    from sdnq.kernels.triton_atten import sdnq_triton_atten
    triton.set_allocator(lambda size, align, stream: torch.empty(size, dtype=torch.int8, device="cuda"))       # device-side TMA descriptors
    shape = (a["batch_per_gpu"], a["heads"], a["tokens"], a["head_dim"])
    q, k, v = (torch.randn(shape, device="cuda").bfloat16() for _ in range(3))
    with torch.no_grad():
        for _ in range(2):
            sdnq_triton_atten(q, k, v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            sdnq_triton_atten(q, k, v)
        e1.record()
        e1.synchronize()
    print("ATTN_REF " + json.dumps({"ms_per_call": e0.elapsed_time(e1) / 5}), flush=True)


def run_attention_workload(args, world, rank, device, peaks):
    """One denoise step's worth of quantized attention through the public entry point (`sdnq_attention`: pre-pass + K9), timed like
    the Linear workloads: CUDA graph of the step, CUDA events, max over ranks; e2e adds the pinned H2D of one call's q / k / v and
    the D2H of the last output.  FLOPs = 4 * batch * heads * tokens^2 * head_dim per call (Q.K^T + P.V)."""
    import torch
    import torch.distributed as dist

    import sdnq_b200
    from sdnq_b200 import _lib
    a = ATTENTION
    shape = (a["batch_per_gpu"], a["heads"], a["tokens"], a["head_dim"])
    flops_step = a["calls"] * 4.0 * shape[0] * shape[1] * shape[2] * shape[2] * shape[3]
    torch.manual_seed(200 + rank)
    host = [torch.randn(shape, dtype=torch.bfloat16).pin_memory() for _ in range(3)]
    # two input sets (2 x 340 MB > 126 MB L2), alternated call by call
    sets = [[h.to(device, non_blocking=True) for h in host], [torch.randn(shape, device=device).bfloat16() for _ in range(3)]]
    holder = {}

    def step():
        for i in range(a["calls"]):
            q, k, v = sets[i & 1]
            holder["y"] = sdnq_b200.sdnq_attention(q, k, v)
        return holder["y"]

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    _lib.launch_count(reset=True)
    step()
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y_graph = step()
    steps, warmup = max(3, min(args.steps, 10)), 3

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        e1.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n

    for _ in range(warmup):
        graph.replay()
    sampler = ClockSampler(device.index) if rank == 0 else None
    if sampler:
        sampler.start()
    t0 = time.time()
    ms = timed(graph.replay, steps)
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None
    out_host = torch.empty(shape, dtype=torch.bfloat16).pin_memory()

    def e2e_step():
        for h, d in zip(host, sets[0]):
            d.copy_(h, non_blocking=True)
        graph.replay()
        out_host.copy_(y_graph, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e2e_step()
    e2e_ms = timed(e2e_step, steps)
    if rank != 0:
        return None
    tf = flops_step * world / ms / 1e9
    res = {"metric": "quantized attention TFLOP/s (4*B*H*N^2*D per call)", "value": tf, "unit": "TFLOP/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8 Q.K^T / bf16 P.V", "data": "synthetic",
           "config": {"workload": a["name"], "describe": a["describe"], "calls_per_step": a["calls"], "batch_per_gpu": a["batch_per_gpu"],
                      "global_batch": a["batch_per_gpu"] * world, "parallelism": f"dp{world}",
                      "l2": "two q/k/v sets of 340 MB alternate call by call (>> 126 MB L2)"},
           "e2e": {"value": flops_step * world / e2e_ms / 1e9, "unit": "TFLOP/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": 3 * host[0].numel() * 2,
                   "d2h_bytes_per_step": out_host.numel() * 2, "copies": "one call's q / k / v in, the last call's output out (the other calls' operands come "
                   "from on-device Linears in a real step)"},
           "gpu_launches": launches_per_step * steps, "launches_per_step": launches_per_step, "clocks": clocks}
    int8_peak, bf16_peak = peaks.get("int8_tflops"), None
    try:
        bf16_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops"))
    except Exception:      # noqa: BLE001
        pass
    if int8_peak and bf16_peak:
        peak = 2.0 / (1.0 / int8_peak + 1.0 / bf16_peak)
        res["roofline"] = {"bound": "tensor", "kernel": "attn_fwd_kernel (K9: tcgen05 kind::i8 Q.K^T + kind::f16 P.V, softmax on CUDA cores / MUFU)",
                           "achieved": flops_step / ms / 1e9, "peak": peak, "unit": "TFLOP/s", "frac": flops_step / ms / 1e9 / peak, "traffic": None,
                           "peak_note": f"harmonic mean of the measured int8 ({int8_peak:.0f}) and bf16 ({bf16_peak:.0f}) dense peaks: half the FLOPs run at each; "
                                        "the kernel's own bound is the softmax (16 MUFU.EX2 per clock and SM = 1024 clocks per 128x128 tile, see DESIGN.md K9); "
                                        "achieved counts the whole call (pre-pass + V transpose + kernel)"}
    if world == 1 and not args.no_gpu_reference:
        ref = {"how": "unmodified reference (oracle/_ref) sdnq_triton_atten on one call's tensors, eager, full autotune space, 5 timed calls after 2 warm-ups; "
                      "the harness installs a triton.set_allocator scratch allocator"}
        try:
            env = dict(os.environ, SDNQ_DEVICE="cuda")
            pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "attention_reference"], capture_output=True, text=True, timeout=300, env=env)
            ln = next((x for x in pr.stdout.splitlines() if x.startswith("ATTN_REF ")), None)
            if ln is None:
                ref["unavailable"] = (pr.stderr.strip().splitlines() or ["no output"])[-1][:300]
            else:
                ms_call = json.loads(ln[9:])["ms_per_call"]
                ref.update(ms_per_call=ms_call, ms_per_step=ms_call * a["calls"], tflops=flops_step / (ms_call * a["calls"]) / 1e9, speedup=ms_call * a["calls"] / ms)
        except Exception as ex:      # noqa: BLE001
            ref["unavailable"] = f"{type(ex).__name__}: {ex}"[:300]
        res["gpu_reference"] = ref
    return res


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from sdnq_b200 import _lib

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
    t_start = time.time()
    peaks = measure_peaks(device) if rank == 0 else {}
    headline = args.workload or HEADLINE
    nested = [] if (args.workload or args.no_nested) else list(NESTED)
    line = run_workload(headline, args, world, rank, device, peaks)
    sub = {}
    for name in nested:
        r = run_workload(name, args, world, rank, device, peaks)
        if r is not None:
            sub[name] = r
    attn = None
    if nested and not args.no_attention:
        try:
            attn = run_attention_workload(args, world, rank, device, peaks)
        except Exception as ex:      # noqa: BLE001  (a nested line must not take the headline down)
            attn = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]} if rank == 0 else None
    if rank == 0:
        line["peaks"] = peaks
        if sub:
            line["workloads"] = sub
        if world == 1 and not args.no_cpu_baseline:
            budget = 20.0
            _, _, info = cpu_reference_run(headline, steps=1000, warmup=1, budget_s=budget)
            line["cpu_baseline"] = info
            for name in sub:
                _, _, sub[name]["cpu_baseline"] = cpu_reference_run(name, steps=1000, warmup=1, budget_s=6.0)
        if world == 1 and not args.no_gpu_reference:
            line["gpu_reference"] = gpu_reference([headline] + list(sub))
        if attn is not None:
            line.setdefault("workloads", {})[ATTENTION["name"]] = attn
        line["bench_seconds"] = round(time.time() - t_start, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, help=f"one of {sorted(WORKLOADS)}; default: {HEADLINE} with {list(NESTED)} nested under 'workloads'")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "gpu_reference", "attention_reference"])
    ap.add_argument("--siblings", default="pool", choices=["off", "block", "pool"],
                    help="grouped launches for sibling projections.  pool (what sdnq_post_load_quant registers): per attention block, and the "
                         "cross-attention to_k / to_v pairs of 4 blocks that read the same encoder states in one launch; block: per block only")
    ap.add_argument("--no-nested", action="store_true")
    ap.add_argument("--no-attention", action="store_true", help="skip the nested quantized-attention line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    args = ap.parse_args()
    if args.impl == "gpu_reference":
        return run_gpu_reference_arm(args)
    if args.impl == "attention_reference":
        return attention_reference_child()
    if args.workload is not None and args.workload not in WORKLOADS:
        ap.error(f"unknown workload {args.workload!r}")
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
