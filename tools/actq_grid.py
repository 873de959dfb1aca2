"""K2 row loop: CTAs per SM of grid (SDNQ_B200_ACTQ_GRID; 0 = one CTA per row block, the round-1 launch shape).  Each setting runs in its
own process (the knob is read once).   python tools/actq_grid.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from sdnq_b200 import ops
from tools.shape_breakdown import graph_time
for (M, K, mode, hg) in [(16384, 3072, "float8_e4m3fn", 256), (18432, 3072, "float8_e4m3fn", 256), (16384, 12288, "float8_e4m3fn", 256), (16384, 3072, "int8", 0),
                         (1024, 1280, "int8", 0), (4096, 640, "int8", 0), (1024, 5120, "int8", 0)]:
    count = max(2, min(16, int(400e6 // (M * K * 2))))
    xs = [torch.randn(M, K, device="cuda", dtype=torch.bfloat16) for _ in range(count)]
    def run():
        for x in xs:
            ops.act_quant(x, mode, hadamard_group=hg)
    t = graph_time(run) / count * 1e3
    print(f"grid={os.environ.get('SDNQ_B200_ACTQ_GRID')}: {M}x{K} {mode} hadamard={hg}: {t:8.2f} us  {3.0 * M * K / t / 1e6:6.2f} TB/s", flush=True)
''' % ROOT
for grid in ("0", "8", "32"):
    env = dict(os.environ, SDNQ_B200_ACTQ_GRID=grid)
    subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
