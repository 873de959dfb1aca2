"""One K9 attention launch for ncu: python tools/attn_one.py Z H QN KN HD [int8|float8_e4m3fn]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdnq_b200 import attention, ops  # noqa: E402

Z, H, QN, KN, HD = (int(a) for a in sys.argv[1:6])
mm = sys.argv[6] if len(sys.argv) > 6 else "int8"
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(Z, H, QN, HD, device="cuda", generator=g).bfloat16()
k = (torch.randn(Z, H, KN, HD, device="cuda", generator=g) + 0.5).bfloat16()
v = torch.randn(Z, H, KN, HD, device="cuda", generator=g).bfloat16()
qq, qs, kq, ks, _, _ = attention.quantize_attn(q, k, v, matmul_dtype=mm)
for _ in range(3):
    out, _ = ops.attention_fwd(qq, kq, v, qs, ks, sm_scale=HD ** -0.5)
torch.cuda.synchronize()
print(float(out.float().abs().mean()))
