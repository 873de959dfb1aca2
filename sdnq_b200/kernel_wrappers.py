"""Lower drop-in boundary: the reference's kernel-level call signatures (kernel_wrappers.py:160-211) bound to the
sm_100a kernels.  `b` is the reference's [K,N] operand, K-major on this platform (stride (1,K)); a contiguous [K,N] `b` is
accepted and re-laid (one copy), exactly as `check_mats` would do upstream (layers/linear/forward.py:9-21)."""
import torch

from . import ops


def _b_nk(b: torch.Tensor) -> torch.Tensor:
    return ops.physical_nk(b.t())


def _mm(a, b, out_dtype):
    out = ops.mm(a.contiguous(), _b_nk(b))
    return out if out_dtype in (None, out.dtype) else out.to(out_dtype)


def int_mm_func(a: torch.Tensor, b: torch.Tensor, out_dtype: torch.dtype = torch.int32) -> torch.Tensor:
    return _mm(a, b, out_dtype)


def fp8_mm_func(a: torch.Tensor, b: torch.Tensor, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    return _mm(a, b, out_dtype)


def _scaled(a, b, scale_a, scale_b, bias, out_dtype):
    sx = scale_a.reshape(-1).to(torch.float32).contiguous()
    sw = scale_b.reshape(-1).to(torch.float32).contiguous()
    return ops.scaled_mm(a.contiguous(), _b_nk(b), sx, sw, bias, out_dtype)


def int_scaled_mm_func(a, b, scale_a, scale_b, bias=None, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    return _scaled(a, b, scale_a, scale_b, bias, out_dtype)


def fp8_scaled_mm_func(a, b, scale_a, scale_b, bias=None, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    return _scaled(a, b, scale_a, scale_b, bias, out_dtype)


def fp_mm_func(a, b, out_dtype=torch.float32):
    raise NotImplementedError("sdnq_b200: the float16 quantized matmul has no sm_100a kernel yet")


fp_scaled_mm_func = fp_mm_func
