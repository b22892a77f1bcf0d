"""Developer probe: a few dcbf_fused launches of one shape (target for ncu).

    python tools/run_shape_once.py A C T M [B] [q8 [sat]]     (DCBF_FLAGS=0x... adds flags to the fused call)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402

A, C, T, M = (int(v) for v in sys.argv[1:5])
q8 = "q8" in sys.argv
if q8:
    sys.argv.remove("q8")
count = "sat" in sys.argv  # q8 with the saturation counter (the slower clamp)
if count:
    sys.argv.remove("sat")
B = int(sys.argv[5]) if len(sys.argv) > 5 else 1
FLAGS = int(os.environ.get("DCBF_FLAGS", "0"), 0)
dev = torch.device("cuda", 0)
x = torch.randint(0, 256, (B, A, C, T, 2, 2), dtype=torch.uint8, device=dev)
dv = torch.rand((C, M, A, 4), dtype=torch.float32, device=dev) * 1e-8
out = torch.empty((B, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
if q8:
    out8 = torch.empty((B, 2, C, T // 16, 16, 2 * M), dtype=torch.int8, device=dev)
    gains = torch.full((M,), 0.01, dtype=torch.float32, device=dev)
    sat = torch.zeros(1, dtype=torch.int64, device=dev)
for _ in range(4):
    if q8:
        _capi.fused_q8(x, dv, gains, out8, B, A, C, C, T, M, 0, 1 / 1712e6, saturated=sat if count else None)
    else:
        _capi.fused(x, dv, out, B, A, C, C, T, M, 0, 1 / 1712e6, FLAGS)
torch.cuda.synchronize()
_capi.fused_status()
print("ok")
