"""Developer probe: a few dcbf_fused_q8 launches at C3 size (target for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402

A, C, T, M, B = 64, 4096, 256, 64, 1
dev = torch.device("cuda", 0)
x = torch.randint(0, 256, (B, A, C, T, 2, 2), dtype=torch.uint8, device=dev)
dv = torch.rand((C, M, A, 4), dtype=torch.float32, device=dev) * 1e-8
out8 = torch.empty((B, 2, C, T // 16, 16, 2 * M), dtype=torch.int8, device=dev)
gains = torch.full((M,), 0.004, dtype=torch.float32, device=dev)
for _ in range(4):
    _capi.fused_q8(x, dv, gains, out8, B, A, C, C, T, M, 0, 1 / 1712e6)
torch.cuda.synchronize()
_capi.fused_status()
print("ok")
