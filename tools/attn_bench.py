"""Quantized attention (K9) timing on the GPU: this library's `sdnq_attention` (pre-pass + kernel) and `attention_fwd` alone, torch SDPA
(bf16, the library yardstick) and -- when oracle/_ref imports and its Triton program compiles on this box -- the reference's own
`sdnq_triton_atten` with its full autotune space.  FLOPs = 4 * Z * H * QN * KN * HD (Q.K^T + P.V).

    python tools/attn_bench.py [--no-reference] [--pv]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = {  # name: Z, H, QN, KN, HD
    "flux_bs4 (24 heads x 4608 tokens, hd 128)": (4, 24, 4608, 4608, 128),
    "flux_bs1": (1, 24, 4608, 4608, 128),
    "sdxl_self_64x64 (10 heads x 4096, hd 64)": (2, 10, 4096, 4096, 64),
    "sdxl_self_32x32 (20 heads x 1024, hd 64)": (2, 20, 1024, 1024, 64),
    "sdxl_cross_32x32 (1024 x 77)": (2, 20, 1024, 77, 64),
}


def timed(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3      # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--pv", action="store_true", help="also time quantised P.V (pv_matmul_dtype = int8 / float8_e4m3fn, Q.K^T int8)")
    ap.add_argument("--shapes", type=int, default=len(SHAPES), help="only the first N shapes")
    args = ap.parse_args()
    import sdnq_b200
    from sdnq_b200 import attention, ops
    ref_fn = None
    if not args.no_reference:
        try:
            from oracle.ref_loader import load_reference
            load_reference(SDNQ_DEVICE="cuda", SDNQ_USE_TORCH_COMPILE="0")
            from sdnq.kernels.triton_atten import sdnq_triton_atten as ref_fn
            import triton
            # device-side TMA descriptors (tl.make_tensor_descriptor) need a scratch allocator from the host program
            triton.set_allocator(lambda size, align, stream: torch.empty(size, dtype=torch.int8, device="cuda"))
        except Exception as e:      # noqa: BLE001
            print(f"reference attention not importable: {type(e).__name__}: {e}")
    for name, (Z, H, QN, KN, HD) in list(SHAPES.items())[:args.shapes]:
        g = torch.Generator(device="cuda").manual_seed(0)
        q = torch.randn(Z, H, QN, HD, device="cuda", generator=g).bfloat16()
        k = (torch.randn(Z, H, KN, HD, device="cuda", generator=g) + 0.5).bfloat16()
        v = torch.randn(Z, H, KN, HD, device="cuda", generator=g).bfloat16()
        flops = 4.0 * Z * H * QN * KN * HD
        line = [f"{name:46s}"]
        for mm in ("int8", "float8_e4m3fn"):
            qq, qs, kq, ks, _, _ = attention.quantize_attn(q, k, v, matmul_dtype=mm)
            t_k = timed(lambda: ops.attention_fwd(qq, kq, v, qs, ks, sm_scale=HD ** -0.5))
            t_e = timed(lambda: sdnq_b200.sdnq_attention(q, k, v, matmul_dtype=mm))
            line.append(f"{mm[:4]}: kernel+V^T {t_k:8.1f} us = {flops / t_k * 1e-6:6.0f} TF/s, with pre-pass {t_e:8.1f} us = {flops / t_e * 1e-6:6.0f} TF/s |")
        if args.pv:
            for pv in ("int8", "float8_e4m3fn"):
                qq, qs, kq, ks, vq, vs = attention.quantize_attn(q, k, v, matmul_dtype="int8", pv_matmul_dtype=pv)
                t_k = timed(lambda: ops.attention_fwd(qq, kq, vq, qs, ks, sm_scale=HD ** -0.5, v_scale=vs))
                t_e = timed(lambda: sdnq_b200.sdnq_attention(q, k, v, matmul_dtype="int8", pv_matmul_dtype=pv))
                line.append(f"P.V {pv[:4]}: kernel+V^T {t_k:8.1f} us = {flops / t_k * 1e-6:6.0f} TF/s, with pre-pass {t_e:8.1f} us = {flops / t_e * 1e-6:6.0f} TF/s |")
        t_s = timed(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
        line.append(f"torch SDPA bf16 {t_s:8.1f} us = {flops / t_s * 1e-6:6.0f} TF/s |")
        if ref_fn is not None:
            try:
                with torch.no_grad():
                    t_r = timed(lambda: ref_fn(q, k, v), iters=5, warmup=2)
                    err = float((ref_fn(q, k, v).float() - sdnq_b200.sdnq_attention(q, k, v).float()).abs().max())
                line.append(f"reference Triton int8 {t_r:8.1f} us = {flops / t_r * 1e-6:6.0f} TF/s (max |diff| to ours {err:.3g})")
            except Exception as e:      # noqa: BLE001
                line.append(f"reference Triton failed: {type(e).__name__}: {str(e)[:200]}")
                ref_fn = None
        print(" ".join(line), flush=True)


if __name__ == "__main__":
    main()
