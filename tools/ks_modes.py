"""Developer probe: K-streamed B tiles vs whole tile sets at the 8-GPU share of C5, for the plain, per-heap-times and int8
modes.    python tools/ks_modes.py
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdk_dc_sand_b200 import _capi
A, C, T, M, B = 197, 512, 256, 256, 1
dev = torch.device("cuda", 0)
x = torch.randint(0, 256, (B, A, C, T, 2, 2), dtype=torch.uint8, device=dev)
dv = torch.rand((C, M, A, 4), dtype=torch.float32, device=dev) * 1e-8
out = torch.empty((B, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
out8 = torch.empty((B, 2, C, T // 16, 16, 2 * M), dtype=torch.int8, device=dev)
gains = torch.full((M,), 0.01, dtype=torch.float32, device=dev)
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); _capi.fused_status()
    return e0.elapsed_time(e1) / n * 1e3
for name, fl in (("k-streamed", 0), ("whole tile sets", _capi.FLAG_DEBUG_NO_KSTREAM)):
    t0 = timed(lambda: _capi.fused(x, dv, out, B, A, C, C, T, M, 0, 1 / 1712e6, fl))
    t1 = timed(lambda: _capi.fused(x, dv, out, B, A, C, C, T, M, 0, 1 / 1712e6, fl, batch_dt=[0.5]))
    t2 = timed(lambda: _capi.fused_q8(x, dv, gains, out8, B, A, C, C, T, M, 0, 1 / 1712e6, fl))
    print(f"C5/8 {name}: plain {t0:.0f} us, per-heap times {t1:.0f} us, int8 output {t2:.0f} us", flush=True)
