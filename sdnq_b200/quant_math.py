"""Load-time quantisation math (PyTorch, runs once per layer on whatever device the weight lives on).

Everything here decides *what is stored*; what is stored is then consumed by the CUDA kernels on every forward.  The
arithmetic (operation order, dtypes, rounding) follows reference quant_utils.py:9-261 so that quantising the same float
weight yields the same integers and scales -- tests/test_host_api.py pins that against reference-generated fixtures."""
import torch

from .common import conv_transpose_types, conv_types, dtype_dict
from .utils import is_pow2, is_pow4, next_power_of_2


# ------------------------------------------------------------------------------------------------ scales
def get_scale_symmetric(weight: torch.Tensor, dim, weights_dtype: str) -> torch.Tensor:
    return weight.abs().amax(dim=dim, keepdim=True).div_(dtype_dict[weights_dtype]["max"])


def get_scale_asymmetric(weight: torch.Tensor, dim, weights_dtype: str):
    info = dtype_dict[weights_dtype]
    if isinstance(dim, int):
        lo, hi = torch.aminmax(weight, dim=dim, keepdim=True)
    else:
        lo, hi = weight.amin(dim=dim, keepdim=True), weight.amax(dim=dim, keepdim=True)
    scale = hi.sub_(lo).div_(info["max"] - info["min"])
    if info["min"] != 0:
        lo.sub_(scale, alpha=info["min"])
    return scale, lo


def quantize_weight(weight: torch.Tensor, dim, weights_dtype: str, dtype: torch.dtype | None = None, use_stochastic_rounding: bool = False):
    """float weight -> (codes in the dtype's torch_dtype, scale, zero_point | None)   (reference quant_utils.py:27-56)."""
    info = dtype_dict[weights_dtype]
    if weight.dtype != torch.float64:
        weight = weight.to(torch.float32, copy=False)
    if info["is_unsigned"]:
        scale, zero_point = get_scale_asymmetric(weight, dim, weights_dtype)
        if dtype is not None:
            scale, zero_point = scale.to(dtype), zero_point.to(dtype)
        q = torch.sub(weight, zero_point).div_(scale)
    else:
        scale, zero_point = get_scale_symmetric(weight, dim, weights_dtype), None
        if dtype is not None:
            scale = scale.to(dtype)
        q = torch.div(weight, scale)
    if info["is_integer"]:
        if use_stochastic_rounding:
            q.add_(torch.randn_like(q), alpha=0.1)
        q.round_()
    else:
        if use_stochastic_rounding:
            step = 1 << (23 - info["mantissa"])
            qi = q.to(torch.float32).view(torch.int32)
            q = qi.add_(torch.randint_like(qi, low=0, high=step, dtype=torch.int32)).bitwise_and_(-step).view(torch.float32)
        q.nan_to_num_()
    q = q.clamp_(info["min"], info["max"]).to(info["torch_dtype"])
    return q, scale, zero_point


def quantize_weight_codebook(weight: torch.Tensor, dim, weights_dtype: str = "uint8", steps: int = 24, dtype: torch.dtype | None = None):
    """Lloyd-Max codebook per reduction group: start from the uniform asymmetric grid, iterate
    assign-to-nearest / recentre, return (indices, sorted levels)   (reference quant_utils.py:59-120)."""
    info = dtype_dict[weights_dtype]
    assert info["is_integer"] and info["is_unsigned"], "Codebook quantization only supports unsigned integer types."
    if weight.dtype != torch.float64:
        weight = weight.to(torch.float32, copy=False)
    shape, ndim = weight.shape, weight.ndim
    flat_all = isinstance(dim, (list, tuple))
    permuted_shape = None
    if flat_all:
        assert len(dim) == ndim, "Codebook quantization only supports quantization along a single dimension."
        weight = weight.flatten()
    else:
        dim = dim + ndim if dim < 0 else dim
        if ndim > 1 and dim != ndim - 1:
            weight = weight.permute(*[i for i in range(ndim) if i != dim], dim)
            permuted_shape = weight.shape
        if ndim > 2:
            weight = weight.flatten(0, -2)
    scatter_dim = 0 if flat_all else 1
    lo, hi = torch.aminmax(weight, dim=-1, keepdim=True)
    step = hi.sub_(lo).div_(info["max"])
    levels = torch.addcmul(lo, torch.arange(info["max"] + 1, dtype=weight.dtype, device=weight.device), step)

    def midpoints(lv):
        return (lv[1:] + lv[:-1]).mul_(0.5) if flat_all else (lv[:, 1:] + lv[:, :-1]).mul_(0.5)

    ones = None
    for _ in range(steps):
        levels = torch.sort(levels, dim=-1).values
        assign = torch.searchsorted(midpoints(levels), weight)
        if ones is None:
            ones = torch.ones_like(assign, dtype=torch.int32)
        counts = torch.zeros_like(levels, dtype=torch.int32).scatter_add_(scatter_dim, assign, ones)
        occupied = counts > 0
        means = torch.zeros_like(levels).scatter_add_(scatter_dim, assign, weight).div_(counts.clamp_(min=1))
        levels = torch.where(occupied, means, levels)
    levels = torch.sort(levels, dim=-1).values
    codes = torch.searchsorted(midpoints(levels), weight, out_int32=True)
    if permuted_shape is not None:
        back = list(range(ndim - 1))
        back.insert(dim, ndim - 1)
        codes = codes.view(permuted_shape).permute(back)
        levels = levels.view(*permuted_shape[:-1], levels.shape[-1]).permute(back)
    else:
        codes = codes.unflatten(0, shape if flat_all else shape[:-1])
        if not flat_all:
            levels = levels.unflatten(0, shape[:-1])
    if dtype is not None:
        levels = levels.to(dtype)
    return codes.to(info["torch_dtype"]), levels


# ------------------------------------------------------------------------------------------------ SVDQuant
def apply_svdquant(weight: torch.Tensor, rank: int = 32, steps: int = 8, dtype: torch.dtype | None = None):
    """W -> (W - up @ down, up [N,r], down [r,K]) with a randomised low-rank SVD   (reference quant_utils.py:123-141)."""
    conv_shape = None
    if weight.ndim > 2:
        conv_shape = weight.shape
        weight = weight.flatten(1, -1)
    if weight.dtype != torch.float64:
        weight = weight.to(torch.float32)
    U, S, V = torch.svd_lowrank(weight, q=rank, niter=steps)
    svd_up = U * S.unsqueeze(0)
    svd_down = V.t_()
    if dtype is not None:
        svd_up, svd_down = svd_up.to(dtype), svd_down.to(dtype)
    residual = weight.sub(torch.mm(svd_up, svd_down))
    if conv_shape is not None:
        residual = residual.unflatten(-1, tuple(conv_shape[1:]))
    return residual, svd_up, svd_down


# ------------------------------------------------------------------------------------------------ layouts for matmul
def prepare_weight_for_matmul(weight: torch.Tensor, matmul_dtype: str | None = "int8") -> torch.Tensor:
    """Give a 2-D B operand [K,N] the K-major layout (stride (1,K)) the tensor-core GEMM reads -- on this platform that is the
    only layout (reference quant_utils.py:239-249 with use_contiguous_*_mm False on cuda)."""
    if weight.is_contiguous():
        weight = weight.t_().contiguous().t_()
    return weight


def prepare_svd_for_matmul(svd_up, svd_down, use_quantized_matmul: bool):
    if svd_up is not None:
        svd_up = prepare_weight_for_matmul(svd_up, "float16") if use_quantized_matmul else svd_up.contiguous()
    if svd_down is not None:
        svd_down = prepare_weight_for_matmul(svd_down, "float16")
    return svd_up, svd_down


# ------------------------------------------------------------------------------------------------ Hadamard (load time)
_H2 = ((1, 1), (1, -1))
_H4 = ((1, 1, 1, -1), (1, 1, -1, 1), (1, -1, 1, 1), (-1, 1, 1, 1))
_HADAMARD_CACHE: dict = {}


def build_hadamard(n: int, dtype: torch.dtype | None = None, device=None) -> torch.Tensor:
    """kron^k(H4)/sqrt(n) when n is a power of 4, Sylvester kron^k(H2)/sqrt(n) otherwise (reference quant_utils.py:144-178)."""
    if not is_pow2(n) or n < 2:
        raise RuntimeError(f"Hadamard Group Size must be a power of 2 but got {n}.")
    base = torch.tensor(_H4 if is_pow4(n) else _H2, dtype=dtype, device=device)
    H = base
    while H.shape[0] < n:
        H = torch.kron(H, base)
    return prepare_weight_for_matmul(H.div_(n ** 0.5), "float16")


def get_hadamard(n: int, dtype: torch.dtype | None = None, device=None) -> torch.Tensor:
    key = (n, torch.device(device) if device is not None else None, dtype)
    H = _HADAMARD_CACHE.get(key)
    if H is None:
        H = _HADAMARD_CACHE[key] = build_hadamard(n, dtype=dtype, device=device)
    return H


def rotate_hadamard(weight: torch.Tensor, group_size: int = 256, hadamard: torch.Tensor | None = None, is_conv: bool = False) -> torch.Tensor:
    """Load-time rotation of a *weight* (x.unflatten(-1,(-1,g)) @ H).  Activations are rotated inside the K2 kernel."""
    if hadamard is None:
        hadamard = get_hadamard(group_size, dtype=weight.dtype, device=weight.device)
    else:
        group_size = hadamard.shape[-1]
        hadamard = hadamard.to(weight.dtype)
    tail = None
    if is_conv:
        tail = list(weight.shape)[1:]
        weight = weight.flatten(1, -1)
    out = torch.matmul(weight.unflatten(-1, (-1, group_size)), hadamard).flatten(-2, -1)
    return out.unflatten(-1, tail) if is_conv else out


def get_hadamard_group_size(channel_size: int, group_size: int):
    g = next_power_of_2(min(channel_size, group_size))
    while channel_size % g != 0:
        g //= 2
    return g >= 4, g


def apply_hadamard(weight: torch.Tensor, group_size: int = 256, hadamard: torch.Tensor | None = None, layer_class_name: str | None = None):
    is_conv = layer_class_name in conv_types or layer_class_name in conv_transpose_types
    if hadamard is not None:
        group_size = hadamard.shape[-1]
    channel_size = weight.shape[1] if is_conv else weight.shape[-1]
    use_hadamard, group_size = get_hadamard_group_size(channel_size, group_size)
    if use_hadamard:
        if hadamard is not None and hadamard.shape[-1] != group_size:
            hadamard = None
        weight = rotate_hadamard(weight, group_size=group_size, hadamard=hadamard, is_conv=is_conv)
    return weight, use_hadamard, group_size
