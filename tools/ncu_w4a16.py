"""One K6 launch at the dominant SD-XL shape for `ncu --set full` (tools/ncu_summary.py full turns the report into profiles/*.md).
    ncu --set full --clock-control none --import-source on -k regex:gemm_w4a16 -c 1 -o out python tools/ncu_w4a16.py [M N K] [nosvd]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdnq_b200 import SDNQConfig, sdnq_quantize_layer

nums = [int(a) for a in sys.argv[1:] if a.isdigit()]
M, N, K = nums if len(nums) == 3 else (1024, 1280, 1280)
svd = "nosvd" not in sys.argv
cfg = dict(weights_dtype="int4", group_size=128, **(dict(use_svd=True, svd_rank=32, svd_steps=2) if svd else {}))
torch.manual_seed(0)
layers = [sdnq_quantize_layer(torch.nn.Linear(K, N, bias=True, device="cuda", dtype=torch.bfloat16), SDNQConfig(**cfg))[0] for _ in range(3)]
x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
for l in layers:
    l(x)
torch.cuda.synchronize()
