"""Quantized attention (row f3), CPU side: the oracle block against plain fp32 softmax attention and its own block-size
independence, and the host mirror's argument handling (no GPU needed: unsupported options must fail before any kernel call)."""
import numpy as np
import pytest
import torch

from oracle import sdnq_oracle as O


def _plain(q, k, v, scale, mask=None):
    s = (q @ k.transpose(0, 1, 3, 2)) * scale
    if mask is not None:
        s = np.where(mask, s, -np.inf)
    p = np.exp(s - s.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    return p @ v


@pytest.mark.parametrize("mm,tol", [("int8", 2e-2), ("float8_e4m3fn", 1.5e-1)])
def test_oracle_attention_is_softmax_attention(mm, tol):
    rng = np.random.default_rng(0)
    Z, H, QN, KN, HD = 1, 2, 70, 90, 64
    q = O.bf16_round(rng.standard_normal((Z, H, QN, HD)).astype(np.float32))
    k = O.bf16_round(rng.standard_normal((Z, H, KN, HD)).astype(np.float32) + 0.5)
    v = O.bf16_round(rng.standard_normal((Z, H, KN, HD)).astype(np.float32))
    qq, qs, kq, ks = O.quantize_attn(q, k, matmul_dtype=mm)
    ref = _plain(q, k, v, HD ** -0.5)
    outs = [O.attn_fwd(qq, kq, v, qs, ks, sm_scale=HD ** -0.5, block_n=bn, out_dtype="float32") for bn in (16, 32, 128)]
    for o in outs:
        assert np.abs(o - ref).max() <= tol * np.abs(ref).max()
    assert np.abs(outs[0] - outs[2]).max() <= 2e-3 * np.abs(ref).max()          # block size only moves roundings
    causal = O.attn_fwd(qq, kq, v, qs, ks, sm_scale=HD ** -0.5, is_causal=True, out_dtype="float32")
    tril = np.tril(np.ones((QN, KN), bool))[None, None]
    masked = O.attn_fwd(qq, kq, v, qs, ks, sm_scale=HD ** -0.5, mask=tril, out_dtype="float32")
    assert np.array_equal(causal, masked)
    add = O.attn_fwd(qq, kq, v, qs, ks, sm_scale=HD ** -0.5, mask=np.where(tril, 0.0, -np.inf).astype(np.float32), out_dtype="float32")
    assert np.abs(add - masked).max() <= 1e-6
    assert np.abs(masked - _plain(q, k, v, HD ** -0.5, tril)).max() <= tol * np.abs(ref).max()


def test_oracle_attention_fully_masked_rows_and_lse():
    rng = np.random.default_rng(1)
    q = rng.standard_normal((1, 1, 5, 16)).astype(np.float32)
    k = rng.standard_normal((1, 1, 40, 16)).astype(np.float32)
    v = rng.standard_normal((1, 1, 40, 16)).astype(np.float32)
    qq, qs, kq, ks = O.quantize_attn(q, k, dtype="float32")
    mask = np.ones((1, 1, 5, 40), bool)
    mask[0, 0, 2] = False
    out, lse = O.attn_fwd(qq, kq, v, qs, ks, mask=mask, sm_scale=0.25, block_n=16, out_dtype="float32", return_lse=True)
    assert np.all(out[0, 0, 2] == 0) and lse[0, 0, 2] == 0            # triton_atten.py:324-331: acc 0 / l_i 1, lse -inf -> 0
    s = ((qq[0, 0] @ kq[0, 0].T) * qs[0, 0, :, None] * ks[0, 0, None, :]) * 0.25 * 1.4426950408889634
    ref_lse = np.log2(np.exp2(s).sum(-1))
    assert np.allclose(np.delete(lse[0, 0], 2), np.delete(ref_lse, 2), rtol=1e-5)


def test_hadamard_group_rule_matches_oracle():
    from sdnq_b200.attention import get_hadamard_group_size
    for channel in (16, 32, 40, 64, 96, 128, 256):
        for group in (4, 64, 128, 256):
            assert get_hadamard_group_size(channel, group) == O.hadamard_group_size(channel, group)


@pytest.mark.parametrize("kwargs", [dict(do_quantize=False), dict(use_fp16_accum=True), dict(pv_matmul_dtype="float16"), dict(matmul_dtype="fp16"),
                                    dict(matmul_dtype="disabled")])
def test_unsupported_attention_options_fail_loudly(kwargs):
    import sdnq_b200
    q = torch.zeros(1, 1, 4, 64, dtype=torch.bfloat16)
    with pytest.raises(NotImplementedError):
        sdnq_b200.sdnq_attention(q, q, q, **kwargs)


def test_attention_needs_a_gpu_tensor():
    import sdnq_b200
    from sdnq_b200 import _lib
    q = torch.zeros(1, 1, 4, 64, dtype=torch.bfloat16)
    with pytest.raises((_lib.SDNQKernelError, RuntimeError)):
        sdnq_b200.sdnq_attention(q, q, q)


def test_oracle_quantised_pv_is_close_to_the_unquantised_result():
    """the P.V branch of the restatement (kernels/triton_atten.py:298-318): int8 / e4m3 codes of p * v_scale with a per-block row scale
    stay within the quantisation noise of the 16-bit P.V, for both key-block sizes used in the GPU tests (the reference's 32, K9's 128)"""
    rng = np.random.default_rng(0)
    Z, H, QN, KN, HD = 1, 2, 70, 200, 32
    q = O.bf16_round(rng.standard_normal((Z, H, QN, HD)).astype(np.float32))
    k = O.bf16_round(rng.standard_normal((Z, H, KN, HD)).astype(np.float32))
    v = O.bf16_round(rng.standard_normal((Z, H, KN, HD)).astype(np.float32))
    qq, qs, kq, ks = O.quantize_attn(q, k, matmul_dtype="int8")
    base = O.attn_fwd(qq, kq, v, qs, ks, sm_scale=HD ** -0.5, out_dtype="float32")
    for pv, tol in (("int8", 2e-2), ("float8_e4m3fn", 6e-2)):
        vq, vs = O.quantize_attn_v(v, pv_matmul_dtype=pv)
        assert vs.shape == (Z, H, KN)
        for bn in (32, 128):
            got = O.attn_fwd(qq, kq, vq, qs, ks, sm_scale=HD ** -0.5, out_dtype="float32", block_n=bn, v_scale=vs, pv_matmul_dtype=pv)
            rel = np.linalg.norm(got - base) / np.linalg.norm(base)
            assert rel <= tol, (pv, bn, rel)


def test_pv_matmul_dtype_aliases_follow_the_reference():
    """triton_atten.py:454-455 ("enabled" / "uint8" -> int8) and :478 (None / "auto" / "none" / "no" / "disabled": unquantised P.V)"""
    from sdnq_b200.attention import _pv_dtype
    for off in (None, "auto", "none", "no", "disabled"):
        assert _pv_dtype(off) is None
    for alias in ("enabled", "uint8", "int8"):
        assert _pv_dtype(alias) == "int8"
    for alias in ("fp8", "float8_e4m3fn"):
        assert _pv_dtype(alias) == "float8_e4m3fn"
    with pytest.raises(NotImplementedError):
        _pv_dtype("float16")
