"""Developer probe: does the fused kernel's time at C3 depend on where its three buffers sit relative to each other?

    python tools/addr_sensitivity.py

One big allocation; samples / delay_vals / beams are carved out of it at different byte offsets (multiples of 2 MiB and
a few sub-page skews) and the kernel is timed for each layout (20 launches between two events).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402

A, C, T, M, B = 64, 4096, 256, 64, 1
dev = torch.device("cuda", 0)
n_x, n_dv, n_out = B * A * C * T * 4, C * M * A * 16, B * 2 * C * T * 2 * M * 4
pool = torch.empty(n_x + n_dv + n_out + (64 << 20), dtype=torch.uint8, device=dev)
pool[: n_x + n_dv].random_(0, 256)


def carve(off, nbytes, dtype, shape):
    return pool[off:off + nbytes].view(dtype).view(shape)


def layout(skew_dv, skew_out):
    off = 0
    x = carve(off, n_x, torch.uint8, (B, A, C, T, 2, 2))
    off += n_x + skew_dv
    dv = carve(off, n_dv, torch.float32, (C, M, A, 4))
    dv.zero_()
    dv[..., 0] = 3e-9
    dv[..., 2] = 0.5
    off += n_dv + skew_out
    out = carve(off, n_out, torch.float32, (B, 2, C, T // 16, 16, 2 * M))
    return x, dv, out


for skew_dv, skew_out in [(0, 0), (2 << 20, 0), (0, 2 << 20), (1 << 20, 3 << 20), (4096, 8192), (65536, 0), (0, 65536),
                          (512 << 10, 256 << 10), (8 << 20, 16 << 20), (1024, 2048)]:
    x, dv, out = layout(skew_dv, skew_out)
    for _ in range(3):
        _capi.fused(x, dv, out, B, A, C, C, T, M, 0, 1 / 1712e6)
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _capi.fused(x, dv, out, B, A, C, C, T, M, 0, 1 / 1712e6)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 20 * 1e3)
    _capi.fused_status()
    print(f"skew dv {skew_dv:>9d} out {skew_out:>9d}: {min(ts):7.1f} us (runs {' '.join(f'{t:.1f}' for t in ts)})", flush=True)
