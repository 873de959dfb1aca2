"""Kernel-level parity: every C-ABI entry point against the numpy oracle and the reference-generated fixtures.
All tests call through sdnq_b200.ops -> ctypes -> libsdnq_b200.so on a CUDA device."""
import numpy as np
import pytest
import torch

from oracle import sdnq_oracle as O
from tests.util import GOLDEN, LAYER_FILES, LAYER_IDS, bf16_ulp_diff, fixture_tensors, np_to_torch, to_f32_np

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ops():
    from sdnq_b200 import ops as _ops
    return _ops


# ----------------------------------------------------------------------------------------------- unpack
@pytest.mark.parametrize("bits", [1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("signed", [False, True])
def test_unpack_int_bit_exact(bits, signed):
    if bits == 1 and signed:
        pytest.skip("int1 is an alias of uint1")
    rng = np.random.default_rng(bits * 2 + signed)
    n = 8 * 4099          # ragged: not a multiple of the block size
    codes = rng.integers(0, 2 ** bits, size=n)
    packed = O.pack_uint(codes, bits)
    name = f"{'' if signed else 'u'}int{bits}"
    expect = codes + (-(2 ** (bits - 1)) if signed else 0)
    tp = torch.from_numpy(packed.astype(np.uint8)).to(DEV)
    for dtype in ((torch.int8, torch.int32, torch.float32) if signed else (torch.uint8, torch.int32, torch.float32)):
        got = ops().unpack(tp, name, (n,), dtype=dtype)
        assert np.array_equal(got.cpu().numpy().astype(np.int64), expect)


def test_unpack_uint1_int64_words():
    rng = np.random.default_rng(5)
    codes = rng.integers(0, 2, size=8 * 1000)
    packed = O.pack_uint(codes, 1).astype(np.int64)      # upstream stores one int64 per packed byte
    got = ops().unpack(torch.from_numpy(packed).to(DEV), "uint1", (codes.size,), dtype=torch.uint8)
    assert np.array_equal(got.cpu().numpy(), codes)


def test_unpack_golden_kat():
    z = np.load(f"{GOLDEN}/pack_kat.npz")
    for bits in range(1, 8):
        codes, packed = z[f"uint{bits}_codes"], z[f"uint{bits}_packed"]
        t = torch.from_numpy(packed.copy()).to(DEV)
        got = ops().unpack(t, f"uint{bits}", codes.shape, dtype=torch.int32)
        assert np.array_equal(got.cpu().numpy(), codes), bits


def test_unpack_minifloat_all_codes():
    z = np.load(f"{GOLDEN}/float_tables.npz")
    for name in [str(n) for n in z["names"]]:
        ref = z[f"{name}_decode"]
        bits = O.dtype_info(name)["num_bits"]
        codes = np.arange(ref.size)
        pad = (-codes.size) % 8
        codes_p = np.concatenate([codes, np.zeros(pad, dtype=codes.dtype)])
        packed = codes_p.astype(np.uint8) if bits == 8 else O.pack_uint(codes_p, bits).astype(np.uint8)
        got = ops().unpack(torch.from_numpy(packed).to(DEV), name, (codes_p.size,), dtype=torch.float32)
        assert np.array_equal(got.cpu().numpy()[: ref.size].view(np.uint32), ref.view(np.uint32)), name


# ----------------------------------------------------------------------------------------------- dequant / requant on fixtures
def _dequant_fixture(t, meta, out_dtype=torch.bfloat16, **kw):
    d = meta["dequantizer"]
    N, K = d["original_shape"]
    return ops().dequant(t["weight"], d["weights_dtype"], t["scale"], t["zero_point"], N, K, d["group_size"], out_dtype,
                         svd_up=t["svd_up"], svd_down=t["svd_down"], svd_layout_matmul=d["use_quantized_matmul"],
                         hadamard_group=d["hadamard_group_size"] if d["use_hadamard"] else 0, use_codebook=d["use_codebook"], **kw)


@pytest.mark.parametrize("path", LAYER_FILES, ids=LAYER_IDS)
def test_dequant_matches_reference(path):
    t, z, meta = fixture_tensors(path, DEV)
    d = meta["dequantizer"]
    W = _dequant_fixture(t, meta)
    Wref = np_to_torch(z["w_dequant"], "bfloat16", DEV)
    assert W.shape == Wref.shape
    du = bf16_ulp_diff(W, Wref)
    if t["svd_up"] is None and not d["use_hadamard"]:
        assert int(du.max()) == 0, f"{int((du > 0).sum())} elements differ, max {int(du.max())} ulp"
    else:
        err = (W.float() - Wref.float()).abs()
        bound = 2.0 ** -8 * Wref.float().abs().amax(dim=-1, keepdim=True)
        assert bool((err <= bound).all()) and float((du > 1).float().mean()) < 1e-3 and float((du > 0).float().mean()) < 0.02


@pytest.mark.parametrize("path", [p for p in LAYER_FILES if "rq_weight__T" in np.load(p).files],
                         ids=[i for p, i in zip(LAYER_FILES, LAYER_IDS) if "rq_weight__T" in np.load(p).files])
def test_requant_matches_reference_bit_exact(path):
    t, z, meta = fixture_tensors(path, DEV)
    d = meta["dequantizer"]
    N, K = d["original_shape"]
    wq, sw, zw, colsum = ops().requant(t["weight"], d["weights_dtype"], t["scale"], t["zero_point"], N, K, d["group_size"],
                                      d["quantized_matmul_dtype"], use_codebook=d["use_codebook"], want_colsum=True)
    ref = z["rq_weight__T"]                                   # physical [N,K]
    got = wq.view(torch.uint8).cpu().numpy() if wq.dtype == torch.float8_e4m3fn else wq.cpu().numpy()
    assert np.array_equal(got.view(ref.dtype), ref)
    assert np.array_equal(sw.cpu().numpy(), z["rq_scale"].reshape(-1))
    if "rq_zero_point" in z.files:
        assert np.array_equal(zw.cpu().numpy(), z["rq_zero_point"].reshape(-1))
    if wq.dtype == torch.int8:
        assert np.array_equal(colsum.cpu().numpy(), ref.astype(np.int64).sum(axis=1))


# ----------------------------------------------------------------------------------------------- activation pre-pass
@pytest.mark.parametrize("mode", ["int8", "uint8", "float8_e4m3fn"])
@pytest.mark.parametrize("M,K", [(1, 64), (33, 640), (77, 2048), (130, 3072), (5, 15360), (64, 256), (19, 1000)])
def test_act_quant_bit_exact(mode, M, K):
    torch.manual_seed(M * 131 + K)
    x = (torch.randn(M, K) * 3).to(torch.bfloat16)
    if M > 4:
        x[2] = 0                         # all-zero row: 0/0 -> code 0, scale 0
        x[3, ::7] *= 50                  # outliers
    xq, sx, zx, rowsum, _ = ops().act_quant(x.to(DEV), mode, want_rowsum=(mode != "float8_e4m3fn"))
    xf = to_f32_np(x)
    if mode == "int8":
        q, s = O.quantize_int_mm(xf)
    elif mode == "uint8":
        q, s, zref = O.quantize_uint_mm(xf)
        assert np.array_equal(zx.cpu().numpy(), zref.reshape(-1))
    else:
        q, s = O.quantize_fp_mm(xf)
        q = O.e4m3fn_bits(q)
    got = xq.view(torch.uint8).cpu().numpy() if mode == "float8_e4m3fn" else xq.cpu().numpy()
    if mode == "float8_e4m3fn":
        mism = got != q
        # +0 / -0 and the zero row are value-equal
        assert np.array_equal(O.from_e4m3fn_bits(got)[mism], O.from_e4m3fn_bits(q)[mism])
    else:
        assert np.array_equal(got, q)
        assert np.array_equal(rowsum.cpu().numpy(), q.astype(np.int64).sum(axis=1))
    assert np.array_equal(sx.cpu().numpy(), s.reshape(-1))


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_act_quant_other_activation_dtypes(dtype):
    torch.manual_seed(3)
    x = torch.randn(50, 1280).to(dtype)
    xq, sx, _, _, _ = ops().act_quant(x.to(DEV), "int8")
    q, s = O.quantize_int_mm(x.float().numpy())
    assert np.array_equal(xq.cpu().numpy(), q) and np.array_equal(sx.cpu().numpy(), s.reshape(-1))


@pytest.mark.parametrize("mode,hg", [("int8", 0), ("uint8", 0), ("float8_e4m3fn", 0), ("int8", 256), ("float8_e4m3fn", 128)])
def test_act_quant_long_rows(mode, hg):
    """K > 16384 (LLM-sized MLPs, wide conv im2col rows): the two-pass kernel; bit-exact codes / scales / row sums."""
    torch.manual_seed(11)
    M, K = 19, 16384 + 2048 + 256
    x = (torch.randn(M, K) * 2).to(torch.bfloat16)
    xq, sx, zx, rowsum, x_rot = ops().act_quant(x.to(DEV), mode, hadamard_group=hg, want_rowsum=True, want_x_rot=True)
    xr = to_f32_np(x_rot)
    if hg:
        ref = O.rotate_hadamard(to_f32_np(x), hg, "bfloat16")
        assert float((bf16_ulp_diff(x_rot.cpu(), torch.from_numpy(ref).to(torch.bfloat16)) > 1).float().mean()) < 2e-3
    else:
        assert torch.equal(x_rot.cpu(), x)
    if mode == "int8":
        q, s = O.quantize_int_mm(xr)
        assert np.array_equal(xq.cpu().numpy(), q) and np.array_equal(sx.cpu().numpy(), s.reshape(-1))
        assert np.array_equal(rowsum.cpu().numpy(), q.astype(np.int32).sum(-1))
    elif mode == "uint8":
        q, s, z = O.quantize_uint_mm(xr)
        assert np.array_equal(xq.cpu().numpy(), q) and np.array_equal(sx.cpu().numpy(), s.reshape(-1)) and np.array_equal(zx.cpu().numpy(), z.reshape(-1))
    else:
        q, s = O.quantize_fp_mm(xr)
        assert np.array_equal(O.from_e4m3fn_bits(xq.view(torch.uint8).cpu().numpy()), np.asarray(q, np.float32))
        assert np.array_equal(sx.cpu().numpy(), s.reshape(-1))


@pytest.mark.parametrize("G", [4, 8, 16, 32, 64, 128, 256])
def test_hadamard_rotation_matches_oracle(G):
    torch.manual_seed(G)
    M, K = 37, 5 * 256 if G == 256 else 640 if G <= 128 and 640 % G == 0 else 512
    x = torch.randn(M, K).to(torch.bfloat16)
    xq, sx, _, _, x_rot = ops().act_quant(x.to(DEV), "int8", hadamard_group=G, want_x_rot=True)
    ref = O.rotate_hadamard(to_f32_np(x), G, "bfloat16")
    du = bf16_ulp_diff(x_rot.cpu(), torch.from_numpy(ref).to(torch.bfloat16))
    err = np.abs(to_f32_np(x_rot) - ref)
    # f32 summation order differs from a GEMM: at most 1 bf16 ulp (or a hair of the row magnitude where sums cancel)
    assert float((du > 1).float().mean()) < 2e-3 and err.max() <= 2.0 ** -7 * np.abs(ref).max()
    assert float((du > 0).float().mean()) < 0.02
    # quantisation of the kernel's own rotated values is exact
    q, s = O.quantize_int_mm(to_f32_np(x_rot))
    assert np.array_equal(xq.cpu().numpy(), q) and np.array_equal(sx.cpu().numpy(), s.reshape(-1))
    # self-inverse: rotating twice returns the input up to bf16 rounding
    _, _, _, _, x_back = ops().act_quant(x_rot, "int8", hadamard_group=G, want_x_rot=True)
    assert float((x_back.float().cpu() - x.float()).abs().max()) <= 0.05 * float(x.float().abs().max())


@pytest.mark.parametrize("G", [8, 16, 32, 64, 128, 256])
@pytest.mark.parametrize("mode", ["int8", "uint8", "float8_e4m3fn"])
def test_hadamard_f16_and_partial_chunks(G, mode):
    """float16 activations (MMA 1 in f16, three bf16 pieces for MMA 2), rows that end in half a chunk, every matmul dtype."""
    torch.manual_seed(100 + G)
    M, K = 45, 256 * 3 + (128 if G <= 128 else 256 * 2)
    x = (torch.randn(M, K) * 3).to(torch.float16)
    xq, sx, zx, rowsum, x_rot = ops().act_quant(x.to(DEV), mode, hadamard_group=G, want_x_rot=True, want_rowsum=True)
    ref = O.rotate_hadamard(x.float().numpy(), G, "float16").astype(np.float16)
    got = x_rot.cpu().numpy()
    ulp = np.abs(got.view(np.int16).astype(np.int32) - ref.view(np.int16).astype(np.int32))
    same_sign = np.signbit(got) == np.signbit(ref)
    assert float(np.mean((ulp > 1) & same_sign)) < 2e-3 and float(np.mean(ulp > 0)) < 0.02
    assert np.abs(got.astype(np.float32) - ref.astype(np.float32)).max() <= 2.0 ** -9 * np.abs(ref.astype(np.float32)).max()
    # quantisation of the kernel's own rotated values is exact
    xr = got.astype(np.float32)
    if mode == "int8":
        q, s = O.quantize_int_mm(xr)
        assert np.array_equal(xq.cpu().numpy(), q) and np.array_equal(sx.cpu().numpy(), s.reshape(-1))
        assert np.array_equal(rowsum.cpu().numpy(), q.astype(np.int32).sum(-1))
    elif mode == "uint8":
        q, s, z = O.quantize_uint_mm(xr)
        assert np.array_equal(xq.cpu().numpy(), q) and np.array_equal(sx.cpu().numpy(), s.reshape(-1))
        assert np.array_equal(zx.cpu().numpy(), z.reshape(-1))
        assert np.array_equal(rowsum.cpu().numpy(), q.astype(np.int32).sum(-1))
    else:
        q, s = O.quantize_fp_mm(xr)
        assert np.array_equal(O.from_e4m3fn_bits(xq.view(torch.uint8).cpu().numpy()), np.asarray(q, np.float32))
        assert np.array_equal(sx.cpu().numpy(), s.reshape(-1))


# ----------------------------------------------------------------------------------------------- GEMM
SHAPES = [(128, 128, 128), (1, 64, 64), (77, 640, 2048), (130, 264, 400), (256, 512, 1024), (300, 1280, 640), (64, 5120, 640),
          (1000, 640, 2560), (33, 48, 16), (513, 776, 208)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_int8_mm_exact(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randint(-128, 128, (M, K), generator=g, dtype=torch.int8)
    b = torch.randint(-128, 128, (N, K), generator=g, dtype=torch.int8)
    got = ops().mm(a.to(DEV), b.to(DEV))
    ref = O.int_mm(a.numpy(), b.numpy().T)
    assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.parametrize("M,N,K", SHAPES[:7])
def test_fp8_mm(M, N, K):
    g = torch.Generator().manual_seed(M + N + K + 1)
    a = (torch.randn(M, K, generator=g) * 2).to(torch.float8_e4m3fn)
    b = (torch.randn(N, K, generator=g) * 2).to(torch.float8_e4m3fn)
    got = ops().mm(a.to(DEV), b.to(DEV))
    ref = O.fp8_mm(a.float().numpy(), b.float().numpy().T)
    # products are exact; only the accumulation order / width of the tensor-core adder differs from f64
    scale = np.abs(a.float().numpy()) @ np.abs(b.float().numpy().T)
    assert np.all(np.abs(got.cpu().numpy() - ref) <= 2e-4 * scale + 1e-6)


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("bias_kind", ["none", "vec", "mat"])
def test_int8_scaled_mm_epilogue(M, N, K, bias_kind):
    g = torch.Generator().manual_seed(M * 3 + N + K)
    a = torch.randint(-128, 128, (M, K), generator=g, dtype=torch.int8)
    b = torch.randint(-128, 128, (N, K), generator=g, dtype=torch.int8)
    sx = torch.rand(M, generator=g) * 0.05 + 1e-3
    sw = torch.rand(N, generator=g) * 0.01 + 1e-4
    bias = None
    if bias_kind == "vec":
        bias = torch.randn(N, generator=g).to(torch.bfloat16)
    elif bias_kind == "mat":
        bias = torch.randn(M, N, generator=g)
    got = ops().scaled_mm(a.to(DEV), b.to(DEV), sx.to(DEV), sw.to(DEV), None if bias is None else bias.to(DEV), torch.bfloat16)
    acc = O.int_mm(a.numpy(), b.numpy().T)
    ref = O.scaled_mm(acc, sx.numpy()[:, None], sw.numpy()[None, :], None if bias is None else bias.float().numpy())
    du = bf16_ulp_diff(got.cpu(), torch.from_numpy(ref).to(torch.bfloat16))
    assert int(du.max()) <= 1 and float((du > 0).float().mean()) < 1e-3
    out32 = ops().scaled_mm(a.to(DEV), b.to(DEV), sx.to(DEV), sw.to(DEV), None if bias is None else bias.to(DEV), torch.float32)
    ref32 = O.scaled_mm(acc, sx.numpy()[:, None], sw.numpy()[None, :], None if bias is None else bias.float().numpy(), out_dtype="float32")
    np.testing.assert_allclose(out32.cpu().numpy(), ref32, rtol=1e-6, atol=1e-6)


def test_scaled_mm_many_tiles_persistent_loop():
    """more tiles than SMs so every CTA walks several tiles and both TMEM accumulator stages are recycled."""
    M, N, K = 2048, 4096, 384
    g = torch.Generator().manual_seed(9)
    a = torch.randint(-128, 128, (M, K), generator=g, dtype=torch.int8)
    b = torch.randint(-128, 128, (N, K), generator=g, dtype=torch.int8)
    got = ops().mm(a.to(DEV), b.to(DEV))
    ref = torch._int_mm(a.to(DEV), b.to(DEV).t()) if hasattr(torch, "_int_mm") else None
    expect = O.int_mm(a.numpy(), b.numpy().T)
    assert np.array_equal(got.cpu().numpy(), expect)
    if ref is not None:
        assert torch.equal(got, ref)


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (300, 520, 400), (1024, 1280, 1280), (513, 776, 208), (4096, 640, 640), (2048, 3072, 3072),
                                   (129, 264, 64), (777, 10240, 1280)])
@pytest.mark.parametrize("kind", ["int8", "fp8", "int8_zp"])
def test_cta_pair_gemm_matches_single_cta(M, N, K, kind, monkeypatch):
    """tcgen05 cta_group::2 kernels (a CTA pair per 256-row tile) against the single-CTA kernels and the exact oracle."""
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randint(-128, 128, (M, K), generator=g, dtype=torch.int8)
    b = torch.randint(-128, 128, (N, K), generator=g, dtype=torch.int8)
    sx = (torch.rand(M, generator=g) * 0.01 + 1e-3).to(DEV)
    sw = (torch.rand(N, generator=g) * 0.01 + 1e-3).to(DEV)
    bias = torch.randn(N, generator=g).to(torch.bfloat16).to(DEV)
    kw = {}
    if kind == "fp8":
        a = (a.float() / 16).to(torch.float8_e4m3fn)
        b = (b.float() / 16).to(torch.float8_e4m3fn)
    if kind == "int8_zp":
        kw = dict(rowsum=a.to(torch.int32).sum(-1).to(torch.int32).to(DEV), zp=torch.randn(N, generator=g).to(DEV))
    a, b = a.to(DEV), b.to(DEV)
    monkeypatch.setenv("SDNQ_B200_CG", "1")
    single = ops().scaled_mm(a, b, sx, sw, bias, torch.bfloat16, **kw)
    for bn in ("128", "256"):
        monkeypatch.setenv("SDNQ_B200_CG", "2")
        monkeypatch.setenv("SDNQ_B200_BN", bn)
        pair = ops().scaled_mm(a, b, sx, sw, bias, torch.bfloat16, **kw)
        monkeypatch.delenv("SDNQ_B200_BN")
        torch.cuda.synchronize()
        if kind == "fp8":
            assert int(bf16_ulp_diff(pair, single).max()) <= 1
        else:
            assert torch.equal(pair, single), f"BN={bn}: max diff {float((pair.float() - single.float()).abs().max())}"
    if kind == "int8":
        acc = O.int_mm(a.cpu().numpy(), b.cpu().numpy().T)
        ref = O.scaled_mm(acc, sx.cpu().numpy().reshape(-1, 1), sw.cpu().numpy().reshape(1, -1), bias.float().cpu().numpy(), out_dtype="bfloat16")
        assert int(bf16_ulp_diff(single.cpu(), torch.from_numpy(ref).to(torch.bfloat16)).max()) <= 1


# ----------------------------------------------------------------------------------------------- K5 small-M Linear (W8A16 GEMV)
@pytest.mark.parametrize("M", [1, 4, 7, 8, 13, 16, 31, 32])
@pytest.mark.parametrize("N,K", [(64, 64), (130, 208), (1280, 1280), (1000, 2816), (18432, 3072)])
@pytest.mark.parametrize("kind", ["int8", "fp8", "int8_zp"])
def test_small_m_linear(M, N, K, kind):
    """y = s[n] * sum_k x q (+ zp[n] * sum_k x) + bias against the same expression in float64, rounded once to bf16: the codes are
    exact in bf16, products exact in f32, accumulation f32 -- so at most a bf16 ulp (or a hair of the row magnitude when sums cancel)."""
    if N * K > 4e6 and M not in (4, 32):
        pytest.skip("large weight: two row counts are enough")
    g = torch.Generator().manual_seed(M * 131 + N + K)
    x = (torch.randn(M, K, generator=g) * 1.5).to(torch.bfloat16)
    q = torch.randint(-128, 128, (N, K), generator=g, dtype=torch.int8)
    if kind == "fp8":
        q = (torch.randn(N, K, generator=g) * 60).clamp(-448, 448).to(torch.float8_e4m3fn)
    sw = torch.rand(N, generator=g) * 0.01 + 1e-3
    zp = torch.randn(N, generator=g) * 0.02 if kind == "int8_zp" else None
    bias = torch.randn(N, generator=g).to(torch.bfloat16)
    y = ops().linear_small_m(x.to(DEV), q.to(DEV), sw.to(DEV), zp=None if zp is None else zp.to(DEV), bias=bias.to(DEV))
    assert y.shape == (M, N) and y.dtype == torch.bfloat16
    xd, qd = x.double(), q.double() if kind != "fp8" else q.float().double()
    ref = (xd @ qd.T) * sw.double() + bias.double()
    mag = (xd.abs() @ qd.abs().T) * sw.double() + bias.double().abs()
    if zp is not None:
        ref = ref + xd.sum(-1, keepdim=True) * zp.double()
        mag = mag + xd.abs().sum(-1, keepdim=True) * zp.double().abs()
    err = (y.cpu().double() - ref).abs()
    # one rounding to bf16 (2^-9 relative) + f32 accumulation error (~K * 2^-24 of the magnitude sum)
    bound = ref.abs() * 2.0 ** -8 + mag * (K * 2.0 ** -23 + 2.0 ** -16)
    assert bool((err <= bound).all()), f"max excess {(err - bound).max()}"
    # f16 activations share the kernel template
    if M == 4 and N <= 1280:
        yh = ops().linear_small_m(x.to(torch.float16).to(DEV), q.to(DEV), sw.to(DEV), zp=None if zp is None else zp.to(DEV), bias=bias.to(torch.float16).to(DEV))
        refh = (x.to(torch.float16).double() @ qd.T) * sw.double() + bias.to(torch.float16).double()
        if zp is not None:
            refh = refh + x.to(torch.float16).double().sum(-1, keepdim=True) * zp.double()
        assert bool(((yh.cpu().double() - refh).abs() <= refh.abs() * 2.0 ** -10 + mag * (K * 2.0 ** -23 + 2.0 ** -16)).all())


def test_small_m_linear_strided_and_errors():
    x = torch.randn(4, 3, 512, device=DEV, dtype=torch.bfloat16)[:, 1]           # rows 1536 elements apart
    q = torch.randint(-128, 128, (96, 512), dtype=torch.int8, device=DEV)
    sw = torch.rand(96, device=DEV) * 0.01
    y = ops().linear_small_m(x, q, sw)
    y2 = ops().linear_small_m(x.contiguous(), q, sw)
    assert torch.equal(y, y2)
    with pytest.raises(ops()._lib.SDNQKernelError):
        ops().linear_small_m(torch.randn(40, 512, device=DEV, dtype=torch.bfloat16), q, sw)      # M > 32
    with pytest.raises(ops()._lib.SDNQKernelError):
        ops().linear_small_m(torch.randn(4, 512, device=DEV), q, sw)                              # f32 activations


# ----------------------------------------------------------------------------------------------- K2 + K1 against the fixtures
@pytest.mark.parametrize("path", [p for p in LAYER_FILES if "mm_xq" in np.load(p).files],
                         ids=[i for p, i in zip(LAYER_FILES, LAYER_IDS) if "mm_xq" in np.load(p).files])
def test_matmul_operands_and_output_match_reference(path):
    t, z, meta = fixture_tensors(path, DEV)
    d = meta["dequantizer"]
    mmdt = d["quantized_matmul_dtype"]
    hg = d["hadamard_group_size"] if d["use_hadamard"] else 0
    xq, sx, zx, rowsum, _ = ops().act_quant(t["x"], mmdt, hadamard_group=hg, want_rowsum=True)
    ref_xq = z["mm_xq"]
    got_xq = xq.view(torch.uint8).cpu().numpy().view(ref_xq.dtype) if xq.dtype == torch.float8_e4m3fn else xq.cpu().numpy()
    if hg:
        assert float(np.mean(got_xq != ref_xq)) < 5e-3
    else:
        if xq.dtype == torch.float8_e4m3fn:
            assert np.array_equal(O.from_e4m3fn_bits(got_xq), O.from_e4m3fn_bits(ref_xq))
        else:
            assert np.array_equal(got_xq, ref_xq)
        assert np.array_equal(sx.cpu().numpy(), z["mm_sx"].reshape(-1))


# ----------------------------------------------------------------------------------------------- rotated 8-bit dequant (tensor-core un-rotate)
@pytest.mark.parametrize("wd", ["int8", "uint8", "float8_e4m3fn"])
@pytest.mark.parametrize("N,K,gs,G", [(96, 1024, -1, 256), (130, 896, 128, 128), (64, 640, 32, 64), (48, 512, -1, 16), (33, 3072, -1, 256)])
def test_dequant_rotated_8bit(wd, N, K, gs, G):
    """use_hadamard layers on the dequant path (dequant_rot8_kernel): scale -> round -> un-rotate, against the oracle."""
    rng = np.random.default_rng(N + K + G)
    groups = K // gs if gs > 0 else 1
    sshape = (N, groups, 1) if groups > 1 else (N, 1)
    scale = (rng.random(sshape) * 0.02 + 0.001).astype(np.float32)
    zp = None
    if wd == "int8":
        w_np = rng.integers(-128, 128, size=(N, K)).astype(np.int8)
        w_t = torch.from_numpy(w_np)
    elif wd == "uint8":
        w_np = rng.integers(0, 256, size=(N, K)).astype(np.uint8)
        w_t = torch.from_numpy(w_np)
        zp = (rng.standard_normal(sshape) * 0.05).astype(np.float32)
    else:
        w_t = torch.from_numpy(np.clip(rng.standard_normal((N, K)).astype(np.float32) * 100, -448, 448)).to(torch.float8_e4m3fn)
        w_np = w_t.float().numpy()
    qshape = [N, groups, gs] if groups > 1 else [N, K]
    layer = O.Layer(w_np.reshape(qshape), scale, zp, weights_dtype=wd, quantized_weight_shape=qshape, result_shape=[N, K] if groups > 1 else None,
                    group_size=gs, use_hadamard=True, hadamard_group_size=G)
    ref = O.dequantize(layer, dtype="bfloat16")
    before = ops()._lib.launch_count(reset=True)
    W = ops().dequant(w_t.to(DEV), wd, torch.from_numpy(scale).to(DEV), None if zp is None else torch.from_numpy(zp).to(DEV), N, K, gs,
                      torch.bfloat16, hadamard_group=G)
    assert ops()._lib.launch_count() == 1
    du = bf16_ulp_diff(W.cpu(), torch.from_numpy(ref).to(torch.bfloat16)).numpy()
    err = np.abs(to_f32_np(W) - ref)
    # f32 summation order differs from a GEMM: one bf16 ulp, except where a group's terms cancel (tiny sums: bound those absolutely)
    solid = np.abs(ref) > 0.05 * np.abs(ref).max(axis=-1, keepdims=True)
    assert float(((du > 1) & solid).mean()) < 1e-3 and float((du > 0).mean()) < 0.03
    assert err.max() <= 2.0 ** -7 * np.abs(ref).max()


# ----------------------------------------------------------------------------------------------- SVD dequant on tensor cores
@pytest.mark.parametrize("wd,bits", [("int4", 4), ("uint4", 4), ("int3", 3), ("uint2", 2)])
@pytest.mark.parametrize("N,K,rank,gs", [(128, 256, 32, 128), (300, 640, 32, 128), (520, 1288, 16, 8), (257, 2048, 64, 64), (1280, 1280, 32, 1280)])
def test_dequant_svd_tensor_core_path(wd, bits, N, K, rank, gs):
    """int<=4-bit + SVD (plain layout) goes through dequant_svd.cu (tcgen05 rank-r update); compare with the oracle."""
    rng = np.random.default_rng(N + K + rank + bits)
    codes = rng.integers(0, 2 ** bits, size=(N, K))
    packed = torch.from_numpy(O.pack_uint(codes, bits).astype(np.uint8))
    groups = K // gs
    sshape = (N, groups, 1) if groups > 1 else (N, 1)
    scale = torch.from_numpy((rng.random(sshape) * 0.02 + 0.001).astype(np.float32))
    zp = torch.from_numpy(rng.standard_normal(sshape).astype(np.float32) * 0.05) if wd.startswith("u") else None
    up = (torch.from_numpy(rng.standard_normal((N, rank)).astype(np.float32)) * 0.1).to(torch.bfloat16)
    down_phys = (torch.from_numpy(rng.standard_normal((K, rank)).astype(np.float32)) * 0.1).to(torch.bfloat16)   # [K,r] = K-major [r,K]
    down = down_phys.t()
    before = ops()._lib.launch_count(reset=True)
    W = ops().dequant(packed.to(DEV), wd, scale.to(DEV), None if zp is None else zp.to(DEV), N, K, gs if groups > 1 else -1, torch.bfloat16,
                      svd_up=up.to(DEV), svd_down=down_phys.to(DEV).t(), svd_layout_matmul=False)
    assert ops()._lib.launch_count() == 1
    layer = O.Layer(packed.numpy(), scale.numpy(), None if zp is None else zp.numpy(), up.float().numpy(), down.float().numpy(),
                    weights_dtype=wd, quantized_weight_shape=[N, groups, gs] if groups > 1 else [N, K], result_shape=[N, K] if groups > 1 else None,
                    group_size=gs if groups > 1 else -1)
    ref = O.dequantize(layer, dtype="bfloat16")
    got = to_f32_np(W)
    du = bf16_ulp_diff(W.cpu(), torch.from_numpy(ref).to(torch.bfloat16))
    bound = 2.0 ** -8 * np.abs(ref).max(axis=-1, keepdims=True)
    assert np.all(np.abs(got - ref) <= bound), float(np.abs(got - ref).max())
    assert float((du > 1).float().mean()) < 1e-3 and float((du > 0).float().mean()) < 0.03


@pytest.mark.parametrize("M,N,K", [(1024, 1280, 1280), (1024, 1280, 5120), (4096, 640, 640), (4096, 640, 2560), (77, 1280, 2048), (300, 1288, 640),
                                   (513, 776, 4112), (128, 128, 38016), (1000, 5120, 640), (256, 256, 128)])
@pytest.mark.parametrize("kind", ["int8", "int8_zp", "uint8", "mat_bias"])
def test_stream_k_gemm_is_bit_identical(M, N, K, kind, monkeypatch):
    """K1 with its k-blocks split evenly over all SMs (parked 32-bit partial accumulators, fixed up by the CTA that holds a tile's first
    k-block) == whole-tile scheduling, bit for bit; repeated launches on one workspace (the flags are re-armed by the kernel)."""
    rng = np.random.default_rng(M + N + K + len(kind))
    a = torch.from_numpy(rng.integers(-128, 128, size=(M, K)).astype(np.int8)).to(DEV)
    b = torch.from_numpy(rng.integers(-128, 128, size=(N, K)).astype(np.int8)).to(DEV)
    sx = torch.from_numpy((rng.random(M) * 0.05 + 1e-3).astype(np.float32)).to(DEV)
    sw = torch.from_numpy((rng.random(N) * 0.01 + 1e-4).astype(np.float32)).to(DEV)
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32)).to(torch.bfloat16).to(DEV)
    kw = {}
    if kind in ("int8_zp", "uint8"):
        kw.update(zp=torch.from_numpy((rng.standard_normal(N) * 0.1).astype(np.float32)).to(DEV), rowsum=a.to(torch.int32).sum(dim=1, dtype=torch.int32))
    if kind == "uint8":
        kw.update(zx=torch.from_numpy((rng.standard_normal(M) * 0.1).astype(np.float32)).to(DEV), colsum=b.to(torch.int32).sum(dim=1, dtype=torch.int32))
    if kind == "mat_bias":
        bias = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32)).to(DEV)
    for out_dtype in (torch.bfloat16, torch.float16):
        monkeypatch.setenv("SDNQ_B200_STREAMK", "0")
        want = ops().scaled_mm(a, b, sx, sw, bias, out_dtype, **kw)
        monkeypatch.setenv("SDNQ_B200_STREAMK", "1")
        for _ in range(3):
            got = ops().scaled_mm(a, b, sx, sw, bias, out_dtype, **kw)
            assert torch.equal(got, want), float((got.float() - want.float()).abs().max())
    # the exact integer accumulator, through the scaled epilogue with unit scales
    ones_m, ones_n = torch.ones(M, device=DEV), torch.ones(N, device=DEV)
    acc = ops().scaled_mm(a[:, :256].contiguous(), b[:, :256].contiguous(), ones_m, ones_n, None, torch.float32)
    if K >= 256:
        assert np.array_equal(acc.cpu().numpy(), O.int_mm(a[:, :256].cpu().numpy(), b[:, :256].cpu().numpy().T).astype(np.float32))


# ----------------------------------------------------------------------------------------------- GEMM with in-kernel int4 unpack
@pytest.mark.parametrize("signed", [True, False])
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (77, 640, 2048), (300, 1288, 640), (1000, 5120, 640), (2048, 4096, 384), (64, 136, 32)])
def test_scaled_mm_packed_int4(signed, M, N, K):
    """packed int4 / uint4 weights expanded by the GEMM's unpack warps == exact integer matmul on the unpacked codes."""
    rng = np.random.default_rng(M + N + K + signed)
    codes = rng.integers(0, 16, size=(N, K))
    packed = torch.from_numpy(O.pack_uint(codes, 4).astype(np.uint8)).to(DEV)
    wvals = codes - 8 if signed else codes
    a = torch.from_numpy(rng.integers(-128, 128, size=(M, K)).astype(np.int8))
    sx = torch.from_numpy((rng.random(M) * 0.05 + 1e-3).astype(np.float32))
    sw = torch.from_numpy((rng.random(N) * 0.01 + 1e-4).astype(np.float32))
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32)).to(torch.bfloat16)
    acc = O.int_mm(a.numpy(), wvals.T.astype(np.int8))
    if signed:
        got = ops().scaled_mm_packed(a.to(DEV), packed, "int4", N, sx.to(DEV), sw.to(DEV), bias.to(DEV), torch.float32)
        ref = O.scaled_mm(acc, sx.numpy()[:, None], sw.numpy()[None, :], bias.float().numpy(), out_dtype="float32")
    else:
        zp = torch.from_numpy(rng.standard_normal(N).astype(np.float32) * 0.1)
        rowsum = torch.from_numpy(a.numpy().astype(np.int64).sum(axis=1).astype(np.int32))
        got = ops().scaled_mm_packed(a.to(DEV), packed, "uint4", N, sx.to(DEV), sw.to(DEV), bias.to(DEV), torch.float32,
                                     rowsum=rowsum.to(DEV), zp=zp.to(DEV))
        zb = (rowsum.numpy().astype(np.float32)[:, None] * sx.numpy()[:, None]).astype(np.float32) * zp.numpy()[None, :]
        zb = (zb.astype(np.float32) + bias.float().numpy()[None, :]).astype(np.float32)
        ref = O.scaled_mm(acc, sx.numpy()[:, None], sw.numpy()[None, :], zb, out_dtype="float32")
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=2e-6, atol=1e-5)


@pytest.mark.parametrize("wd", ["int2", "uint2", "int3", "uint3", "int5", "uint5", "int6", "uint6", "int7", "uint7", "uint4", "float6_e3m2fn", "float7_e3m3fn",
                                "float5_e2m2fn", "float4_e2m1fn", "float7_e4m2fn", "float6_e2m3fn", "float3_e1m1fn", "float5_e3m2fnu"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (77, 640, 2048), (300, 1288, 640), (1000, 384, 1280)])
def test_scaled_mm_packed_any_width(wd, M, N, K):
    """2..7-bit integer / minifloat weights expanded by the GEMM's unpack warps == the unpack kernel's operand through the plain GEMM,
    bit for bit (same codes, same accumulator, same epilogue)."""
    from sdnq_b200.common import dtype_dict
    info = dtype_dict.get(wd)
    if info is None:
        pytest.skip(f"{wd} is not a storage format of this build")
    bits = info["num_bits"]
    if (K * bits) % 128 != 0:
        pytest.skip("row pitch of the packed rows is not a multiple of 16 bytes")
    rng = np.random.default_rng(M + N + K + bits)
    codes = rng.integers(0, 2 ** bits, size=(N, K))
    packed = torch.from_numpy(O.pack_uint(codes, bits).astype(np.uint8)).to(DEV)
    sx = torch.from_numpy((rng.random(M) * 0.05 + 1e-3).astype(np.float32)).to(DEV)
    sw = torch.from_numpy((rng.random(N) * 0.01 + 1e-4).astype(np.float32)).to(DEV)
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32)).to(torch.bfloat16).to(DEV)
    kw = {}
    if info["is_integer"]:
        a = torch.from_numpy(rng.integers(-128, 128, size=(M, K)).astype(np.int8)).to(DEV)
        b = ops().unpack(packed, wd, (N, K), dtype=torch.int8)
        if info["is_unsigned"]:
            kw = dict(zp=torch.from_numpy((rng.standard_normal(N) * 0.1).astype(np.float32)).to(DEV), rowsum=a.to(torch.int32).sum(dim=1, dtype=torch.int32))
        else:
            assert int(b.min()) == -(2 ** (bits - 1)) and int(b.max()) == 2 ** (bits - 1) - 1
    else:
        a = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float32)).to(torch.float8_e4m3fn).to(DEV)
        b = ops().unpack(packed, wd, (N, K), dtype=torch.float8_e4m3fn)
        assert torch.equal(b.float(), ops().unpack(packed, wd, (N, K), dtype=torch.float32))       # the format is a subset of e4m3
    for out_dtype in (torch.bfloat16, torch.float32):
        want = ops().scaled_mm(a, b, sx, sw, bias, out_dtype, **kw)
        got = ops().scaled_mm_packed(a, packed, wd, N, sx, sw, bias, out_dtype, **kw)
        assert torch.equal(got, want), float((got.float() - want.float()).abs().max())


# ----------------------------------------------------------------------------------------------- load-time quantisation (K8)
@pytest.mark.parametrize("wd", ["int8", "uint8", "int7", "uint7", "int6", "uint6", "int5", "uint5", "int4", "uint4", "int3", "uint3", "int2", "uint2"])
@pytest.mark.parametrize("N,K,gs", [(64, 256, 32), (33, 640, 128), (16, 1024, -1), (5, 2048, 256), (7, 96, 8), (3, 4104, -1), (9, 1536, 512), (2, 64, 16)])
def test_quantize_weight_matches_host_arithmetic(wd, N, K, gs):
    """scale + round + clamp + pack in one kernel == quant_math.quantize_weight + packing.pack_int on the CPU (the reference's
    arithmetic, pinned to its fixtures by tests/test_host_api.py): same packed bytes, same scales, same zero points."""
    from sdnq_b200 import packing, quant_math
    from sdnq_b200.common import dtype_dict
    info = dtype_dict[wd]
    g = torch.Generator().manual_seed(N * 7 + K + info["num_bits"])
    for wdtype, scale_dtype in ((torch.float32, None), (torch.bfloat16, None), (torch.float32, torch.bfloat16)):
        w = (torch.randn(N, K, generator=g) * torch.rand(N, 1, generator=g) * 3).to(wdtype)
        w[0, : (K if gs <= 0 else gs)] = 0                                        # an all-zero group: 0 / 0 -> code of 0
        w[-1, -1] = 1e4                                                           # an outlier
        groups = 1 if gs <= 0 else K // gs
        view = w.float().view(N, groups, K // groups)
        q, s_ref, z_ref = quant_math.quantize_weight(view, -1, wd, dtype=scale_dtype)
        want = packing.pack_int(q, wd).reshape(-1).view(torch.uint8) if info["is_packed"] else q.reshape(-1).view(torch.uint8)
        codes, scale, zp = ops().quantize_weight(w.to(DEV), wd, gs, scale_dtype)
        assert np.array_equal(scale.cpu().numpy().reshape(-1), s_ref.float().numpy().reshape(-1)), "scale"
        if info["is_unsigned"]:
            assert np.array_equal(zp.cpu().numpy().reshape(-1), z_ref.float().numpy().reshape(-1)), "zero point"
        else:
            assert zp is None and z_ref is None
        assert np.array_equal(codes.cpu().reshape(-1).view(torch.uint8).numpy(), want.numpy()), "codes"


FLOAT_WD = ["float8_e4m3fn", "float8_e5m2", "float8_e4m3fn_sdnq", "float8_e5m2fn", "float8_e3m4fn", "float8_e4m4fnu", "float7_e3m3fn", "float7_e4m2fn",
            "float7_e2m5fnu", "float6_e3m2fn", "float6_e2m3fn", "float6_e1m4fn", "float5_e2m2fn", "float5_e4m0fn", "float4_e2m1fn", "float4_e3m0fn",
            "float4_e2m2fnu", "float3_e2m0fn", "float3_e1m1fn", "float2_e1m0fn"]


@pytest.mark.parametrize("wd", FLOAT_WD)
@pytest.mark.parametrize("N,K,gs", [(64, 256, 32), (33, 640, 128), (16, 1024, -1), (7, 96, 8), (3, 4104, -1), (9, 1536, 512)])
def test_quantize_weight_float_formats_match_host_arithmetic(wd, N, K, gs):
    """the float formats through K8: scale + divide + nan_to_num + clamp + (cast to torch.float8_* | pack_float's bit arithmetic) + pack
    == quant_math.quantize_weight + packing.pack_float on the CPU (packed_float.py:26-82, pinned to the reference's float fixtures by
    tests/test_host_api.py): same bytes, same scales, same zero points"""
    from sdnq_b200 import packing, quant_math
    from sdnq_b200.common import dtype_dict
    info = dtype_dict[wd]
    g = torch.Generator().manual_seed(N * 5 + K + info["num_bits"] + info["mantissa"])
    for wdtype, scale_dtype in ((torch.float32, None), (torch.bfloat16, None), (torch.float32, torch.bfloat16)):
        w = (torch.randn(N, K, generator=g) * torch.rand(N, 1, generator=g) * 3).to(wdtype)
        w[0, : (K if gs <= 0 else gs)] = 0                                        # an all-zero group: 0 / 0 -> 0
        w[-1, -1] = 1e4                                                           # an outlier: the rest of its group lands in the subnormals
        w[1, :8] = torch.tensor([1e-30, -1e-30, 0.0, -0.0, 1e-3, -1e-3, 2.5e-1, -2.5e-1]).to(wdtype)
        groups = 1 if gs <= 0 else K // gs
        view = w.float().view(N, groups, K // groups)
        q, s_ref, z_ref = quant_math.quantize_weight(view, -1, wd, dtype=scale_dtype)
        want = packing.pack_float(q, wd).reshape(-1).view(torch.uint8) if info["is_packed"] else q.reshape(-1).view(torch.uint8)
        codes, scale, zp = ops().quantize_weight(w.to(DEV), wd, gs, scale_dtype)
        assert np.array_equal(scale.cpu().numpy().reshape(-1), s_ref.float().numpy().reshape(-1)), "scale"
        if info["is_unsigned"]:
            assert np.array_equal(zp.cpu().numpy().reshape(-1), z_ref.float().numpy().reshape(-1)), "zero point"
        else:
            assert zp is None and z_ref is None
        got = codes.cpu().reshape(-1).view(torch.uint8).numpy()
        assert np.array_equal(got, want.numpy()), f"codes: {int((got != want.numpy()).sum())} bytes differ"


@pytest.mark.parametrize("path", [p for p in LAYER_FILES if "small_m" not in p], ids=[i for p, i in zip(LAYER_FILES, LAYER_IDS) if "small_m" not in p])
def test_quantize_weight_reproduces_reference_fixture(path):
    """the stored `weight` / `scale` / `zero_point` bytes of a layer the reference quantised, from its float weight"""
    t, z, meta = fixture_tensors(path)
    d = meta["dequantizer"]
    from sdnq_b200.common import dtype_dict
    info = dtype_dict[d["weights_dtype"]]
    if (info["num_bits"] < 2 or info["num_bits"] > 8 or d.get("use_hadamard") or d.get("use_codebook") or t["svd_up"] is not None
            or d["group_size"] == -2):
        pytest.skip("outside the quantisation kernel (2..8-bit formats without rotation / SVD / codebook)")
    N, K = meta["N"], meta["K"]
    gs = d["group_size"]
    codes, scale, zp = ops().quantize_weight(t["w_orig"].to(DEV), d["weights_dtype"], gs if gs > 0 else -1, None if t["scale"].dtype == torch.float32 else t["scale"].dtype)
    stored = t["weight"]
    stored = ops().physical_nk(stored) if stored.ndim == 2 and not info["is_packed"] else stored
    assert np.array_equal(codes.cpu().reshape(-1).view(torch.uint8).numpy(), stored.contiguous().reshape(-1).view(torch.uint8).numpy())
    assert np.array_equal(scale.cpu().numpy().reshape(-1), t["scale"].float().numpy().reshape(-1))
    if t["zero_point"] is not None:
        assert np.array_equal(zp.cpu().numpy().reshape(-1), t["zero_point"].float().numpy().reshape(-1))


# ----------------------------------------------------------------------------------------------- SVD branch of the W8A8 forwards (K7 + rank-r accumulate in K1)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,K,r", [(1, 16, 8), (77, 2048, 32), (130, 640, 16), (1024, 1280, 32), (333, 1296, 64), (40, 3072, 24), (16, 48, 32)])
def test_svd_low(dtype, M, K, r):
    """K7: low = cast(x @ svd_down^T) with f32 accumulation -- torch.mm's rounding point (linear_int8.py:59)."""
    g = torch.Generator().manual_seed(M + K + r)
    x = torch.randn(M, K, generator=g).to(dtype)
    down = (torch.randn(r, K, generator=g) / K ** 0.5).to(dtype)
    got = ops().svd_low(x.to(DEV), down.to(DEV)).cpu()
    ref32 = x.double() @ down.double().t()
    ref = ref32.to(dtype)
    assert got.dtype == dtype and got.shape == (M, r)
    # f32 accumulation in a different order than the f64 reference: the rounded results agree to one unit in the last place
    err = (got.double() - ref32).abs()
    ulp = torch.maximum(ref32.abs(), torch.tensor(1e-3, dtype=torch.float64)) * (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10)
    assert bool((err <= ulp).all()), float((err / ulp).max())
    assert float((got != ref).float().mean()) < 0.02
    # strided rows
    xs = torch.randn(M, K + 8, generator=g).to(dtype).to(DEV)[:, :K]
    got2 = ops().svd_low(xs, down.to(DEV)).cpu()
    ref2 = (xs.cpu().double() @ down.double().t())
    assert bool(((got2.double() - ref2).abs() <= torch.maximum(ref2.abs(), torch.tensor(1e-3, dtype=torch.float64)) * (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10)).all())


@pytest.mark.parametrize("kind", ["int8", "fp8", "uint8", "int4", "uint4"])
@pytest.mark.parametrize("M,N,K,r", [(128, 128, 128, 32), (77, 640, 2048, 32), (300, 1288, 640, 16), (1000, 5120, 640, 64), (2048, 4096, 384, 32), (64, 136, 32, 32)])
def test_scaled_mm_svd(kind, M, N, K, r):
    """K1 with the rank-r second accumulate: out = fma(acc * sx, sw, zero-point terms + low @ svd_up^T + bias), the sum formed in f32."""
    rng = np.random.default_rng(M + N + K + r + len(kind))
    dt = torch.bfloat16
    low = torch.from_numpy(rng.standard_normal((M, r)).astype(np.float32)).to(dt)
    up = torch.from_numpy((rng.standard_normal((N, r)) / r ** 0.5).astype(np.float32)).to(dt)
    sx = torch.from_numpy((rng.random(M) * 0.05 + 1e-3).astype(np.float32))
    sw = torch.from_numpy((rng.random(N) * 0.01 + 1e-4).astype(np.float32))
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32)).to(dt)
    svd32 = (low.double() @ up.double().t()).numpy()
    kw = {}
    extra = np.zeros((M, N), dtype=np.float64)
    if kind in ("int4", "uint4"):
        codes = rng.integers(0, 16, size=(N, K))
        b = torch.from_numpy(O.pack_uint(codes, 4).astype(np.uint8))
        wvals = codes - 8 if kind == "int4" else codes
        a = torch.from_numpy(rng.integers(-128, 128, size=(M, K)).astype(np.int8))
        acc = O.int_mm(a.numpy(), wvals.T.astype(np.int8)).astype(np.float64)
        kw = dict(packed_dtype=kind, N=N)
        if kind == "uint4":
            zp = torch.from_numpy((rng.standard_normal(N) * 0.1).astype(np.float32))
            rowsum = torch.from_numpy(a.numpy().astype(np.int64).sum(axis=1).astype(np.int32))
            kw.update(zp=zp.to(DEV), rowsum=rowsum.to(DEV))
            extra = rowsum.numpy().astype(np.float64)[:, None] * sx.numpy().astype(np.float64)[:, None] * zp.numpy().astype(np.float64)[None, :]
    elif kind == "fp8":
        a = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float32)).to(torch.float8_e4m3fn)
        b = torch.from_numpy(rng.standard_normal((N, K)).astype(np.float32)).to(torch.float8_e4m3fn)
        acc = (a.double() @ b.double().t()).numpy()
    else:
        a = torch.from_numpy(rng.integers(-128, 128, size=(M, K)).astype(np.int8))
        b = torch.from_numpy(rng.integers(-128, 128, size=(N, K)).astype(np.int8))
        acc = O.int_mm(a.numpy(), b.numpy().T).astype(np.float64)
        if kind == "uint8":            # asymmetric activations and weights: zero-point, column-sum and K * zx * zp terms
            zp = torch.from_numpy((rng.standard_normal(N) * 0.1).astype(np.float32))
            zx = torch.from_numpy((rng.standard_normal(M) * 0.1).astype(np.float32))
            rowsum = torch.from_numpy(a.numpy().astype(np.int64).sum(axis=1).astype(np.int32))
            colsum = torch.from_numpy(b.numpy().astype(np.int64).sum(axis=1).astype(np.int32))
            kw.update(zp=zp.to(DEV), rowsum=rowsum.to(DEV), zx=zx.to(DEV), colsum=colsum.to(DEV))
            f = lambda t: t.numpy().astype(np.float64)
            extra = (f(rowsum)[:, None] * f(sx)[:, None] * f(zp)[None, :] + f(colsum)[None, :] * f(sw)[None, :] * f(zx)[:, None]
                     + K * f(zx)[:, None] * f(zp)[None, :])
    got = ops().scaled_mm_svd(a.to(DEV), b.to(DEV), sx.to(DEV), sw.to(DEV), low.to(DEV), up.to(DEV), bias.to(DEV), dt, **kw).float().cpu().numpy()
    ref = acc * sx.numpy().astype(np.float64)[:, None] * sw.numpy().astype(np.float64)[None, :] + extra + svd32 + bias.double().numpy()[None, :]
    mag = np.abs(acc * sx.numpy()[:, None] * sw.numpy()[None, :]) + np.abs(extra) + np.abs(svd32) + np.abs(bias.float().numpy())[None, :]
    err = np.abs(got - ref)
    bound = np.maximum(np.abs(ref), 1e-3) * 2.0 ** -8 + mag * 2e-6 + (mag * 3e-3 if kind == "fp8" else 0)   # half a bf16 ulp + f32 accumulation slack
    assert np.all(err <= bound), float((err / bound).max())
    # without a bias the SVD product alone is the addend
    got_nb = ops().scaled_mm_svd(a.to(DEV), b.to(DEV), sx.to(DEV), sw.to(DEV), low.to(DEV), up.to(DEV), None, dt, **kw).float().cpu().numpy()
    ref_nb = ref - bias.double().numpy()[None, :]
    assert np.all(np.abs(got_nb - ref_nb) <= np.maximum(np.abs(ref_nb), 1e-3) * 2.0 ** -8 + mag * 2e-6 + (mag * 3e-3 if kind == "fp8" else 0))


# ----------------------------------------------------------------------------------------------- grouped launch (sibling projections)
@pytest.mark.parametrize("kind", ["int8", "int8_zp", "fp8", "uint8", "int4", "uint4"])
@pytest.mark.parametrize("M,K,ns,align", [(1024, 1280, (1280, 1280, 1280), 256), (4096, 640, (640, 640, 640), 128), (77, 2048, (1280, 1280), 256),
                                          (77, 2048, (640, 640), 128), (300, 384, (136, 1288, 8, 520), 128), (2048, 256, (3072, 3072, 3072, 12288), 256),
                                          (130, 512, (256,) * 8, 256), (33, 64, (512, 256), 128)])
def test_scaled_mm_grouped_equals_separate_launches(kind, M, K, ns, align):
    """one grouped launch over the stacked sibling operands == every sibling's own scaled_mm, bit for bit."""
    rng = np.random.default_rng(M + K + sum(ns) + len(kind))
    packed = kind in ("int4", "uint4")
    if packed and align == 256:
        align = 128
    starts = [0]
    for n in ns:
        starts.append(starts[-1] + (n + align - 1) // align * align)
    Nt = starts[-1]
    sx = torch.from_numpy((rng.random(M) * 0.05 + 1e-3).astype(np.float32)).to(DEV)
    sw = torch.zeros(Nt)
    bias = torch.zeros(Nt)
    zp = torch.zeros(Nt)
    colsum = torch.zeros(Nt, dtype=torch.int32)
    if kind == "fp8":
        a = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float32)).to(torch.float8_e4m3fn).to(DEV)
        b = torch.zeros((Nt, K), dtype=torch.float8_e4m3fn)
    elif packed:
        a = torch.from_numpy(rng.integers(-128, 128, size=(M, K)).astype(np.int8)).to(DEV)
        b = torch.zeros((Nt, K // 2), dtype=torch.uint8)
    else:
        a = torch.from_numpy(rng.integers(-128, 128, size=(M, K)).astype(np.int8)).to(DEV)
        b = torch.zeros((Nt, K), dtype=torch.int8)
    for s, n in zip(starts, ns):
        if kind == "fp8":
            b[s:s + n] = torch.from_numpy(rng.standard_normal((n, K)).astype(np.float32)).to(torch.float8_e4m3fn)
        elif packed:
            b[s:s + n] = torch.from_numpy(rng.integers(0, 256, size=(n, K // 2)).astype(np.uint8))
        else:
            b[s:s + n] = torch.from_numpy(rng.integers(-128, 128, size=(n, K)).astype(np.int8))
            colsum[s:s + n] = b[s:s + n].to(torch.int32).sum(dim=1, dtype=torch.int32)
        sw[s:s + n] = torch.from_numpy((rng.random(n) * 0.01 + 1e-4).astype(np.float32))
        bias[s:s + n] = torch.from_numpy(rng.standard_normal(n).astype(np.float32))
        zp[s:s + n] = torch.from_numpy((rng.standard_normal(n) * 0.1).astype(np.float32))
    b, sw, bias, zp, colsum = b.to(DEV), sw.to(DEV), bias.to(DEV), zp.to(DEV), colsum.to(DEV)
    kw = {}
    if kind in ("int8_zp", "uint8", "uint4"):
        kw.update(zp=zp, rowsum=a.to(torch.int32).sum(dim=1, dtype=torch.int32))
    if kind == "uint8":
        kw.update(zx=torch.from_numpy((rng.standard_normal(M) * 0.1).astype(np.float32)).to(DEV), colsum=colsum)
    pk = kind if packed else None
    for out_dtype in (torch.bfloat16, torch.float32):
        outs = ops().scaled_mm_grouped(a, b, sx, sw, starts, list(ns), bias, out_dtype, packed_dtype=pk, **kw)
        assert len(outs) == len(ns)
        for g, (s, n) in enumerate(zip(starts, ns)):
            sub = {k: (v[s:s + n] if k in ("zp", "colsum") else v) for k, v in kw.items()}
            if packed:
                sub.pop("colsum", None)
                one = ops().scaled_mm_packed(a, b[s:s + n].contiguous(), kind, n, sx, sw[s:s + n], bias[s:s + n], out_dtype, **sub)
            else:
                one = ops().scaled_mm(a, b[s:s + n].contiguous(), sx, sw[s:s + n], bias[s:s + n], out_dtype, **sub)
            assert outs[g].shape == (M, n) and outs[g].is_contiguous()
            assert torch.equal(outs[g], one), (g, float((outs[g].float() - one.float()).abs().max()))


# ----------------------------------------------------------------------------------------------- fused quantise + GEMM (one launch)
FUSED_SHAPES = [(1024, 1280, 1280), (4096, 640, 640), (77, 1280, 2048), (333, 136, 272), (1, 64, 32), (32, 8, 16), (128 * 5 + 7, 648, 5120),
                (2500, 256, 4096 + 16), (148 * 2 + 1, 384, 16384)]


@pytest.mark.parametrize("mm", ["int8", "float8_e4m3fn"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", FUSED_SHAPES)
def test_linear_w8a8_fused_matches_two_kernel_path(mm, dtype, M, N, K):
    """linear_w8a8(fused=True) (single launch: in-kernel activation quantiser + cross-CTA strip flags) must be bit-identical to
    act_quant followed by scaled_mm, launch after launch on the same workspace (the counters re-arm themselves)."""
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    x = (torch.randn(M, K, generator=g) * 3).to(dtype).to(DEV)
    if M > 2:
        x[1].zero_()                                     # all-zero row: scale 0, codes 0
        x[2, 0] = 1e4                                    # outlier row
    if mm == "int8":
        w = torch.randint(-127, 128, (N, K), generator=g, dtype=torch.int8).to(DEV)
    else:
        w = (torch.randn(N, K, generator=g) * 2).to(torch.float8_e4m3fn).to(DEV)
    sw = (torch.rand(N, generator=g) * 0.01 + 1e-4).to(DEV)
    bias = torch.randn(N, generator=g).to(dtype).to(DEV)
    o = ops()
    o._lib.launch_count(reset=True)
    got = o.linear_w8a8(x, w, mm, sw, bias=bias, out_dtype=dtype, fused=True)
    assert o._lib.launch_count() == 1, "expected the single fused launch"
    xq, sx, *_ = o.act_quant(x, mm)
    ref = o.scaled_mm(xq, w, sx, sw, bias, dtype)
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    for _ in range(3):                                   # same workspace again, and without bias
        got2 = o.linear_w8a8(x, w, mm, sw, bias=None, out_dtype=dtype, fused=True)
    ref2 = o.scaled_mm(xq, w, sx, sw, None, dtype)
    assert torch.equal(got2.view(torch.int16), ref2.view(torch.int16))
    torch.cuda.synchronize()
    ws = next(iter(o._WORKSPACES.values()))
    assert int(ws[:4096].view(torch.int32).abs().sum()) == 0, "strip counters must be left at zero"


def test_linear_w8a8_fused_strided_rows_and_graph_replay():
    """x with a row stride (ldx > K), interleaved shapes on one workspace, and CUDA-graph replay with changing inputs."""
    o = ops()
    g = torch.Generator(device="cpu").manual_seed(5)
    big = torch.randn(600, 1024, generator=g).to(torch.bfloat16).to(DEV)
    x = big[:, 128:128 + 640]                            # ldx = 1024, K = 640, 16 B aligned
    w = torch.randint(-127, 128, (328, 640), generator=g, dtype=torch.int8).to(DEV)
    w2 = torch.randint(-127, 128, (640, 1024), generator=g, dtype=torch.int8).to(DEV)
    sw = (torch.rand(328, generator=g) * 0.01 + 1e-4).to(DEV)
    sw2 = (torch.rand(640, generator=g) * 0.01 + 1e-4).to(DEV)

    def both():
        return (o.linear_w8a8(x, w, "int8", sw, out_dtype=torch.bfloat16, fused=True),
                o.linear_w8a8(big, w2, "int8", sw2, out_dtype=torch.bfloat16, fused=True))

    def refs():
        xq, sx, *_ = o.act_quant(x, "int8")
        bq, bs, *_ = o.act_quant(big, "int8")
        return o.scaled_mm(xq, w, sx, sw, None, torch.bfloat16), o.scaled_mm(bq, w2, bs, sw2, None, torch.bfloat16)

    for a, b in zip(both(), refs()):
        assert torch.equal(a.view(torch.int16), b.view(torch.int16))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        both()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ya, yb = both()
    for seed in (1, 2, 3):
        big.copy_(torch.randn(600, 1024, generator=torch.Generator().manual_seed(seed)).to(torch.bfloat16))
        graph.replay()
        torch.cuda.synchronize()
        for a, b in zip((ya, yb), refs()):
            assert torch.equal(a.view(torch.int16), b.view(torch.int16))


# ----------------------------------------------------------------------------------------------- flat dequant kernel (int4 / int8)
@pytest.mark.parametrize("wd", ["int4", "uint4", "int8", "uint8"])
@pytest.mark.parametrize("N,K,gs", [(64, 640, 128), (33, 1280, 32), (7, 24, -1), (100, 1296, 8), (5, 16, 16), (257, 2048, -1), (1280, 5120, 128), (3, 4104, -1)])
@pytest.mark.parametrize("out_dtype", ["bfloat16", "float32"])
def test_dequant_flat_path_bit_exact(wd, N, K, gs, out_dtype):
    """int4 / int8 weights with in-row groups (or row-wise scales) take the flat, fully coalesced kernel; bit-exact vs the oracle."""
    bits = 4 if wd.endswith("4") else 8
    rng = np.random.default_rng(N * 3 + K + bits)
    codes = rng.integers(0, 2 ** bits, size=(N, K))
    if bits == 4:
        stored = O.pack_uint(codes, 4).astype(np.uint8)
    else:
        stored = codes.astype(np.uint8) if wd == "uint8" else (codes - 128).astype(np.int8)
    groups = K // gs if gs > 0 else 1
    sshape = (N, groups, 1) if groups > 1 else (N, 1)
    scale = (rng.random(sshape) * 0.02 + 0.001).astype(np.float32)
    zp = (rng.standard_normal(sshape) * 0.05).astype(np.float32) if wd.startswith("u") else None
    td = getattr(torch, out_dtype)
    w_dev = torch.from_numpy(stored).to(DEV)
    ops()._lib.launch_count(reset=True)
    W = ops().dequant(w_dev, wd, torch.from_numpy(scale).to(DEV), None if zp is None else torch.from_numpy(zp).to(DEV), N, K,
                      gs if groups > 1 else -1, td)
    assert ops()._lib.launch_count() == 1
    layer = O.Layer(stored if bits == 4 else (stored.reshape(N, groups, K // groups) if groups > 1 else stored), scale, zp, None, None, weights_dtype=wd,
                    quantized_weight_shape=[N, groups, gs] if groups > 1 else [N, K], result_shape=[N, K] if groups > 1 else None,
                    group_size=gs if groups > 1 else -1)
    ref = O.dequantize(layer, dtype=out_dtype)
    assert np.array_equal(to_f32_np(W).reshape(N, K), np.asarray(ref, dtype=np.float32).reshape(N, K))
