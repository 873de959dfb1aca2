"""Attention operand pre-pass on the GPU, kernel by kernel (FLUX 4 x 24 x 4608 x 128 by default): the row quantiser alone (q), the
channel means + row quantiser (smooth-K k), and the bytes each moves against the copy peak.

    python tools/attn_prepass_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


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
    return a.elapsed_time(b) / iters * 1e3


def main():
    from sdnq_b200 import _lib, ops
    Z, H, N, HD = 4, 24, 4608, 128
    x = torch.randn(Z, H, N, HD, device="cuda").bfloat16()
    nbytes = x.numel() * 2
    lib = _lib.load()
    mean = torch.empty((Z * H, HD), dtype=torch.float32, device="cuda")
    t_q = timed(lambda: ops.attn_quant(x, "int8"))
    t_k = timed(lambda: ops.attn_quant(x, "int8", smooth=True))
    t_m = timed(lambda: _lib.check(lib.sdnq_b200_attn_colmean(x.data_ptr(), ops.dtype_code(x.dtype), Z * H, N, HD, mean.data_ptr(), torch.cuda.current_stream().cuda_stream)))
    t_c = timed(lambda: x.clone())
    print(f"row quantiser            {t_q:7.1f} us  {1.5 * nbytes / t_q * 1e-6:6.0f} GB/s (reads {nbytes / 1e6:.0f} MB, writes {nbytes / 2e6:.0f} MB)")
    print(f"channel means            {t_m:7.1f} us  {nbytes / t_m * 1e-6:6.0f} GB/s (reads {nbytes / 1e6:.0f} MB)")
    print(f"means + row quantiser    {t_k:7.1f} us")
    print(f"torch clone (copy peak)  {t_c:7.1f} us  {2 * nbytes / t_c * 1e-6:6.0f} GB/s")


if __name__ == "__main__":
    main()
