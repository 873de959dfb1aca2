"""torch.library registration of the kernels (`sdnq_b200::*`) so they can sit inside torch.compile / export graphs.

The reference registers its Triton kernels as `triton_op("sdnq::scaled_mm")` (kernels/triton_scaled_mm.py:239) for the same reason.
Each op is a thin wrapper over `sdnq_b200.ops` (the C-ABI call) with a fake implementation that only describes the outputs, so
tracing needs no GPU.  The eager forwards in `forward.py` call `ops` directly (no dispatcher overhead); use these from compiled code:

    torch.ops.sdnq_b200.scaled_mm(xq, wq_nk, sx, sw, bias, out_dtype)
    torch.ops.sdnq_b200.act_quant(x, "int8", hadamard_group)
    torch.ops.sdnq_b200.dequant_rowwise(w, "int8", scale, out_dtype)
    torch.ops.sdnq_b200.linear_small_m(x, wq_nk, sw, zp, bias)
"""
import torch

from . import ops

_MM_TORCH = {"int8": torch.int8, "uint8": torch.int8, "float8_e4m3fn": torch.float8_e4m3fn, "fp8": torch.float8_e4m3fn}


@torch.library.custom_op("sdnq_b200::scaled_mm", mutates_args=())
def scaled_mm(a: torch.Tensor, b_nk: torch.Tensor, sx: torch.Tensor, sw: torch.Tensor, bias: torch.Tensor | None,
              out_dtype: torch.dtype) -> torch.Tensor:
    """K1: out[M,N] = cast(fma(f32(a @ b_nk^T) * sx[m], sw[n], bias)); a [M,K] and b_nk [N,K] int8 or float8_e4m3fn."""
    return ops.scaled_mm(a, b_nk, sx, sw, bias, out_dtype)


@scaled_mm.register_fake
def _(a, b_nk, sx, sw, bias, out_dtype):
    return a.new_empty((a.shape[0], b_nk.shape[0]), dtype=out_dtype)


@torch.library.custom_op("sdnq_b200::act_quant", mutates_args=())
def act_quant(x: torch.Tensor, matmul_dtype: str, hadamard_group: int) -> tuple[torch.Tensor, torch.Tensor]:
    """K2 (symmetric modes): x [..., K] -> (codes [M,K], row scales [M])."""
    xq, sx, _, _, _ = ops.act_quant(x, matmul_dtype, hadamard_group=hadamard_group)
    return xq, sx


@act_quant.register_fake
def _(x, matmul_dtype, hadamard_group):
    m = x.numel() // x.shape[-1]
    return x.new_empty((m, x.shape[-1]), dtype=_MM_TORCH[matmul_dtype]), x.new_empty((m,), dtype=torch.float32)


@torch.library.custom_op("sdnq_b200::dequant_rowwise", mutates_args=())
def dequant_rowwise(weight: torch.Tensor, weights_dtype: str, scale: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    """K3 for a row-wise 8-bit weight [N,K] (the stored matmul operand): W = cast(q * scale[n])."""
    w = ops.physical_nk(weight)
    return ops.dequant(w, weights_dtype, scale, None, w.shape[0], w.shape[1], -1, out_dtype)


@dequant_rowwise.register_fake
def _(weight, weights_dtype, scale, out_dtype):
    n, k = (weight.shape if weight.is_contiguous() else weight.t().shape)
    return weight.new_empty((n, k), dtype=out_dtype)


@torch.library.custom_op("sdnq_b200::linear_small_m", mutates_args=())
def linear_small_m(x: torch.Tensor, wq_nk: torch.Tensor, sw: torch.Tensor, zp: torch.Tensor | None, bias: torch.Tensor | None) -> torch.Tensor:
    """K5: fewer than 33 rows of x times the 1-byte weight codes, read once."""
    return ops.linear_small_m(x, wq_nk, sw, zp=zp, bias=bias)


@linear_small_m.register_fake
def _(x, wq_nk, sw, zp, bias):
    return x.new_empty((*x.shape[:-1], wq_nk.shape[0]))
