"""Developer probe: same-box A/B of dcbf_fused across flag words (debug bits included) on one shape.

    python tools/time_flags.py A C T M flags [flags ...]      (interleaved rounds, CUDA events, 20 launches each)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402

A, C, T, M = (int(v) for v in sys.argv[1:5])
FLAGS = [int(v, 0) for v in sys.argv[5:]] or [0]
dev = torch.device("cuda", 0)
x = torch.randint(0, 256, (1, A, C, T, 2, 2), dtype=torch.uint8, device=dev)
dv = torch.rand((C, M, A, 4), dtype=torch.float32, device=dev) * 1e-8
out = torch.empty((1, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
byts = _capi.load().dcbf_fused_bytes(1, A, C, T, M)
for f in FLAGS:
    for _ in range(5):
        _capi.fused(x, dv, out, 1, A, C, C, T, M, 0, 1 / 1712e6, f)
torch.cuda.synchronize()
for rnd in range(4):
    for f in FLAGS:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _capi.fused(x, dv, out, 1, A, C, C, T, M, 0, 1 / 1712e6, f)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"round {rnd} flags {f:#8x}: {us:8.1f} us  {byts / us / 1e3 / 6550.1:.3f} of 6550 GB/s", flush=True)
_capi.fused_status()
