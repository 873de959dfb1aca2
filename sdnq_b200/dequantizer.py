"""SDNQDequantizer: per-layer metadata + the two weight-side entry points of the forward.

  __call__            -> K3 kernel (unpack + scale [+zp | codebook] [+svd] [un-rotate] -> W[N,K])   reference dequantizer.py:389-429
  re_quantize_matmul  -> K4 kernel (unpack + dequant f32 + row-wise re-quantise to int8 / fp8)     reference dequantizer.py:353-386

Same dataclass fields as the reference (dequantizer.py:279-350) because `loader` / `apply_sdnq_options_to_model` and
serialised checkpoints address them by name.  Difference by design: the reference recomputes the re-quantised operand on
every forward; weights are frozen (requires_grad=False), so here it is computed once per layer and cached on the module
(see forward._matmul_operand)."""
from dataclasses import dataclass

import torch

from . import ops
from .common import conv_transpose_types, conv_types, dtype_dict


@dataclass
class SDNQDequantizer:
    result_dtype: torch.dtype
    result_shape: torch.Size
    original_shape: torch.Size
    original_stride: list
    quantized_weight_shape: torch.Size
    weights_dtype: str
    quantized_matmul_dtype: str
    hadamard_group_size: int
    group_size: int
    svd_rank: int
    svd_steps: int
    codebook_steps: int
    use_quantized_matmul: bool
    re_quantize_for_matmul: bool
    use_stochastic_rounding: bool
    use_hadamard: bool
    use_codebook: bool
    layer_class_name: str

    def __post_init__(self):
        w, m = dtype_dict[self.weights_dtype], dtype_dict[self.quantized_matmul_dtype]
        self.num_bits, self.is_packed, self.is_integer, self.is_unsigned = w["num_bits"], w["is_packed"], w["is_integer"], w["is_unsigned"]
        self.num_bits_matmul, self.is_packed_matmul = m["num_bits"], m["is_packed"]
        self.is_integer_matmul, self.is_unsigned_matmul = m["is_integer"], m["is_unsigned"]

    # ---- geometry helpers --------------------------------------------------------------------
    @property
    def is_conv(self) -> bool:
        return self.layer_class_name in conv_types or self.layer_class_name in conv_transpose_types

    def matmul_nk(self):
        """(N, K) of the GEMM a W8A8 layer runs: Linear [N,K]; Conv [N, C*kh*kw] (the im2col contraction)."""
        shape = tuple(self.original_shape)
        if self.layer_class_name in conv_types:
            k = 1
            for v in shape[1:]:
                k *= int(v)
            return int(shape[0]), k
        return self._linear_nk()

    def _linear_nk(self):
        if self.is_conv:
            if self.layer_class_name in conv_types and self.use_quantized_matmul:
                return self.matmul_nk()        # the flattened [N, C*kh*kw] weight: same GEMM geometry as a Linear
            raise NotImplementedError(f"sdnq_b200: {self.layer_class_name} weights have no quantized-matmul operand (the reference "
                                      "runs ConvTranspose layers on the dequant path only)")
        shape = tuple(self.original_shape)
        if len(shape) != 2:
            raise NotImplementedError(f"sdnq_b200: only 2-D Linear weights have a CUDA dequant kernel (got shape {shape})")
        return shape

    # ---- K3 ------------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, weight, scale, zero_point=None, svd_up=None, svd_down=None, hadamard=None, skip_quantized_matmul: bool = False,
                 non_hadamard: bool = False, skip_compile: bool = False, dtype: torch.dtype | None = None) -> torch.Tensor:
        """Dequantised weight [N,K] in `dtype` (default result_dtype).  `skip_quantized_matmul=True` means the tensors are stored in
        matmul layout (K-major weight / transposed SVD factors); physically that is the same [N,K] weight, so only the SVD strides
        differ.  `hadamard` (a matrix in the reference) is accepted for signature compatibility; the kernel un-rotates with its
        own butterfly."""
        dtype = dtype or self.result_dtype
        ops._require_cuda(weight)          # raises: sdnq_b200 has no CPU dequant path
        if self.is_conv:
            return self._conv_dequant(weight, scale, zero_point, svd_up, svd_down, skip_quantized_matmul, non_hadamard, dtype)
        N, K = self._linear_nk()
        un_rotate = self.hadamard_group_size if (self.use_hadamard and not non_hadamard) else 0
        return ops.dequant(weight, self.weights_dtype, scale, zero_point, N, K, self.group_size, dtype, svd_up=svd_up, svd_down=svd_down,
                           svd_layout_matmul=bool(skip_quantized_matmul), hadamard_group=un_rotate, use_codebook=self.use_codebook)

    def _conv_dequant(self, weight, scale, zero_point, svd_up, svd_down, skip_quantized_matmul, non_hadamard, dtype):
        """Conv / ConvTranspose weights: broadcast dequant over the quantised view (reference dequantizer.py:15-84 with
        `is_conv`): scale [N,1,kh,kw] / grouped [N,C/g,1,kh,kw] / ConvTranspose [1,N,kh,kw] ...; the SVD term is
        mm(svd_up, svd_down) in the SVD dtype added in f32 (no intermediate rounding of the weight, unlike Linear)."""
        stored_t = bool(skip_quantized_matmul and not self.re_quantize_for_matmul and self.use_quantized_matmul)
        un_rotate = self.use_hadamard and not non_hadamard
        if un_rotate and stored_t:
            # dequantize_symmetric (dequantizer.py:66, 82-83) decides `is_conv` from the *stored* tensor's rank: a weight kept in
            # matmul layout is 2-D, so the reference rotates the 4-D result along its last (kernel-width) axis and fails in unflatten
            raise NotImplementedError("sdnq_b200: a Hadamard-rotated convolution weight stored in matmul layout cannot be dequantised "
                                      "(the reference raises on this combination too: rotate_hadamard over the kernel-width axis)")
        if stored_t:      # matmul layout: weight [K,N] K-major (physically [N,K]), scale / zp [1,N]
            N, K = self.matmul_nk()
            weight = ops.physical_nk(weight)
            view = (N, K)
            scale = scale.reshape(N, 1)
            zero_point = None if zero_point is None else zero_point.reshape(N, 1)
        else:
            view = tuple(self.quantized_weight_shape)
        addend = None
        if svd_up is not None:
            if skip_quantized_matmul:      # the factors are transposed whenever the layer is in matmul mode, re-quantised or not
                svd_up, svd_down = svd_up.t().contiguous(), svd_down.contiguous().t()      # (quantizer.py:164-167, dequantizer.py:30-36)
            addend = torch.mm(svd_up, svd_down)
        out = ops.dequant_nd(weight, self.weights_dtype, scale, zero_point, view, dtype, use_codebook=self.use_codebook, addend=addend)
        shape = tuple(self.result_shape) if self.result_shape is not None else tuple(self.original_shape)
        if un_rotate:
            # rotate_hadamard(result, is_conv=True) (quant_utils.py:193-209): groups of hadamard_group_size along the flattened
            # [N, C*kh*kw] weight, product rounded to the result dtype -- the rotation stage of K2 on the dense matrix
            flat = out.view(shape[0], -1)
            out = ops.act_quant(flat, "int8", hadamard_group=self.hadamard_group_size, want_x_rot=True)[4]
        return out.view(shape)

    # ---- K4 ------------------------------------------------------------------------------------
    @torch.no_grad()
    def re_quantize_matmul(self, weight, scale, zero_point=None, svd_up=None, svd_down=None, hadamard=None, non_hadamard: bool = True,
                           skip_compile: bool = False):
        """-> (Wq [K,N] K-major view, sw [1,N][, zp [1,N]]) exactly as the reference returns them (dequantizer.py:166-239).
        SVD / Hadamard are never applied here (the reference passes none on this path)."""
        wq, sw, zw, _ = self.re_quantize_matmul_raw(weight, scale, zero_point)
        if self.is_integer_matmul and self.is_unsigned_matmul:
            return wq.t(), sw.unsqueeze(0), zw.unsqueeze(0)
        return wq.t(), sw.unsqueeze(0)

    @torch.no_grad()
    def re_quantize_matmul_raw(self, weight, scale, zero_point=None, want_colsum: bool = False):
        """physical form for the kernels: (wq [N,K], sw [N], zw [N] | None, colsum [N] | None)"""
        N, K = self._linear_nk()
        if self.is_conv:
            # re_quantize_*_mm on a convolution (dequantizer.py:166-200): dequantise over the quantised view in f32 (K3c), flatten to
            # [N, C*kh*kw], row-wise quantise -- the row quantiser is K2 run on the dense f32 matrix (same arithmetic, bit-exact).
            # Run once per layer; forward.matmul_operand caches the result.
            dense = ops.dequant_nd(weight, self.weights_dtype, scale, zero_point, tuple(self.quantized_weight_shape), torch.float32,
                                   use_codebook=self.use_codebook).view(N, K)
            uint8_mm = self.is_integer_matmul and self.is_unsigned_matmul
            wq, sw, zw, colsum, _ = ops.act_quant(dense, self.quantized_matmul_dtype, want_rowsum=want_colsum or uint8_mm)
            return wq, sw, zw, colsum
        return ops.requant(weight, self.weights_dtype, scale, zero_point, N, K, self.group_size, self.quantized_matmul_dtype,
                           use_codebook=self.use_codebook, want_colsum=want_colsum)


torch.serialization.add_safe_globals([SDNQDequantizer])
