"""Developer probe: dcbf_fused_ex with static, per-heap and per-tile time-varying steering and with beam weights on the
BASELINE shapes (CUDA events, 20 launches each).  DCBF_LIB=/path/to/other/libdcbf.so for same-box A/B runs.

    python tools/time_steering_modes.py
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdk_dc_sand_b200 import _capi
def run(A, C, T, M, B, tv, sample_dt=0.0, w=False):
    dev = torch.device("cuda", 0)
    x = torch.randint(0, 256, (B, A, C, T, 2, 2), dtype=torch.uint8, device=dev)
    dv = torch.rand((C, M, A, 4), dtype=torch.float32, device=dev) * 1e-8
    out = torch.empty((B, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
    wt = torch.rand((M, A), dtype=torch.float32, device=dev) if w else None
    dt = [0.1 * i for i in range(B)] if tv else None
    f = lambda: _capi.fused_ex(x, dv, out, B, A, C, C, T, M, 0, 1 / 1712e6, batch_dt=dt, sample_dt=sample_dt, weights=wt)
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize(); _capi.fused_status()
    us = e0.elapsed_time(e1) / 20 * 1e3
    byts = _capi.load().dcbf_fused_bytes(B, A, C, T, M)
    print(f"A{A} C{C} T{T} M{M} B{B} tv={tv} sample_dt={sample_dt!r} weights={w}: {us:.1f} us  {byts/us/1e3/6550.1:.3f}", flush=True)
for shape in ((64, 4096, 256, 64, 1), (80, 4096, 256, 32, 1), (197, 512, 256, 256, 1), (64, 1024, 256, 16, 1)):
    run(*shape, False); run(*shape, True); run(*shape, True, sample_dt=4.8e-6); run(*shape, False, w=True)
