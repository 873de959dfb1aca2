"""Generate the committed golden fixtures of the quantized attention row by running the UNMODIFIED
reference -- its own `get_attn_inputs` / `quantize_attn` and its Triton `sdnq_attn_kernel` -- on a GPU box:

    gpurun -- 'python tests/golden/generate_attention.py gpurun_out/attention_golden.npz'     # then copy it to tests/golden/

(the reference's attention is a Triton program: it cannot run in the authoring container, which has no GPU; oracle/_ref -- the scripted
byte-for-byte copy of the reference package -- travels to the box with the snapshot).  The autotune space is pinned to one point
(BLOCK_SIZE_M = 128, BLOCK_SIZE_N = 32, 4 warps, 2 stages) so that the quantised-P.V cases, whose result depends on the key-block
size, are reproducible.  Everything written is data produced by running the reference; bf16 tensors are stored as uint16 bit
patterns, fp8 codes as uint8 bit patterns.  `tests/test_oracle_golden.py` holds the numpy oracle against these on the CPU."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.update(SDNQ_TRITON_ATTEN_BLOCK_SIZE_M_LIST="128", SDNQ_TRITON_ATTEN_BLOCK_SIZE_N_LIST="32", SDNQ_TRITON_ATTEN_NUM_WARPS_LIST="4",
                  SDNQ_TRITON_ATTEN_NUM_STAGES_LIST="2")
from oracle.ref_loader import load_reference  # noqa: E402

sdnq = load_reference(SDNQ_DEVICE="cuda", SDNQ_USE_TORCH_COMPILE="0")
import triton  # noqa: E402
from sdnq.kernels import triton_atten  # noqa: E402

# the reference builds its TMA descriptors on the device, which needs a scratch allocator from the host program (harness set-up)
triton.set_allocator(lambda size, align, stream: torch.empty(size, dtype=torch.int8, device="cuda"))

CASES = {  # name: (Z, H, KH, QN, KN, HD, kwargs of sdnq_triton_atten)
    "int8": (1, 2, 2, 160, 96, 64, {}),
    "fp8": (1, 2, 2, 96, 160, 64, dict(matmul_dtype="float8_e4m3fn")),
    "causal": (1, 1, 1, 128, 128, 64, dict(is_causal=True)),
    "bool_mask": (2, 2, 2, 64, 96, 64, dict(mask="bool")),
    "additive_mask": (1, 2, 2, 64, 64, 64, dict(mask="float")),
    "gqa_no_smooth": (1, 4, 2, 64, 128, 64, dict(smooth_k=False)),
    "hd128": (1, 1, 1, 128, 64, 128, {}),
    "pv_int8": (1, 2, 2, 96, 128, 64, dict(pv_matmul_dtype="int8")),
    "pv_fp8": (1, 1, 1, 64, 96, 64, dict(pv_matmul_dtype="float8_e4m3fn")),
}


def bits(t):
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16)
    if t.dtype in (torch.float8_e4m3fn, torch.float8_e5m2):
        return t.view(torch.uint8).numpy()
    if t.dtype == torch.bool:
        return t.numpy().astype(np.uint8)
    return t.numpy()


def main(out_path):
    arrays, meta = {}, {"block_size_n": 32, "cases": {}}
    for name, (Z, H, KH, QN, KN, HD, kw) in CASES.items():
        g = torch.Generator().manual_seed(1000 + len(name) * 7 + QN + KN)
        q = torch.randn(Z, H, QN, HD, generator=g).bfloat16().cuda()
        k = (torch.randn(Z, KH, KN, HD, generator=g) + 0.5 * torch.randn(1, KH, 1, HD, generator=g)).bfloat16().cuda()
        v = torch.randn(Z, KH, KN, HD, generator=g).bfloat16().cuda()
        kw = dict(kw)
        mask_kind = kw.pop("mask", None)
        mask = None
        if mask_kind == "bool":
            mask = torch.rand(Z, 1, QN, KN, generator=g) > 0.3
            mask[..., 0] = True
            mask[0, 0, 5] = False                       # a fully masked row
            mask = mask.cuda()
        elif mask_kind == "float":
            mask = torch.randn(1, H, QN, KN, generator=g).cuda()
        if mask is not None:
            kw["attn_mask"] = mask
        with torch.no_grad():
            qq, qs, kq, ks, vq, vs, _, _ = triton_atten.quantize_attn(q, k, v, smooth_k=kw.get("smooth_k", True), matmul_dtype=kw.get("matmul_dtype", "int8"),
                                                                  pv_matmul_dtype=kw.get("pv_matmul_dtype"))
            out = triton_atten.sdnq_triton_atten(q, k, v, **kw)
        torch.cuda.synchronize()
        for key, t in (("q", q), ("k", k), ("v", v), ("q_codes", qq), ("q_scale", qs), ("k_codes", kq), ("k_scale", ks), ("out", out)):
            arrays[f"{name}.{key}"] = bits(t)
        if vs is not None:
            arrays[f"{name}.v_codes"], arrays[f"{name}.v_scale"] = bits(vq), bits(vs)
        if mask is not None:
            arrays[f"{name}.mask"] = bits(mask)
        meta["cases"][name] = dict(Z=Z, H=H, KH=KH, QN=QN, KN=KN, HD=HD, mask=mask_kind, out_dtype=str(out.dtype).replace("torch.", ""),
                                   code_dtype=str(qq.dtype).replace("torch.", ""),
                                   kwargs={k_: v_ for k_, v_ in kw.items() if k_ != "attn_mask"})
    meta["torch"], meta["triton"], meta["device"] = torch.__version__, triton.__version__, torch.cuda.get_device_name(0)
    np.savez_compressed(out_path, meta=np.array(json.dumps(meta)), **arrays)
    print(f"wrote {out_path}: {len(arrays)} arrays, {os.path.getsize(out_path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "attention_golden.npz"))
