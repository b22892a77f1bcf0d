import os, sys
import torch
sys.path.insert(0, "/root/repo")
from dpdk_dc_sand_b200 import _capi
TS = 1 / 1712e6
dev = torch.device("cuda", 0)
bad = 0
for (A, C, T, M, B, flags) in [(64, 4096, 256, 64, 1, 0), (64, 300, 256, 64, 2, 0), (64, 170, 256, 16, 1, 0), (80, 40, 256, 32, 1, 0), (5, 7, 48, 3, 2, 0), (33, 200, 144, 8, 1, 2), (1, 20, 64, 16, 1, 1)]:
    x = torch.randint(0, 256, (B, A, C, T, 2, 2), dtype=torch.uint8, device=dev)
    dv = torch.zeros((C, M, A, 4), dtype=torch.float32, device=dev)
    dv[..., 0] = (torch.rand((C, M, A), device=dev) * 32 - 16) * TS
    dv[..., 2] = (torch.rand((C, M, A), device=dev) * 2 - 1) * 3.14159265
    gains = (torch.rand(M, device=dev) * 0.02 + 0.002).float()
    want = torch.zeros((B, 2, C, T // 16, 16, 2 * M), dtype=torch.int8, device=dev)
    got = torch.full_like(want, 77)
    sat_w = torch.zeros(1, dtype=torch.int64, device=dev); sat_g = torch.zeros(1, dtype=torch.int64, device=dev)
    packed = torch.empty(_capi.fused_packed_bytes(A, C, M, flags), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    _capi.fused_q8(x, dv, gains, want, B, A, C, 2 * C, T, M, 1, TS, flags, saturated=sat_w)
    _capi.fused_pack_coeffs_q8(dv, gains, packed, A, C, 2 * C, M, 1, TS, flags & 2)
    _capi.fused_packed_q8(x, packed, gains, got, B, A, C, 2 * C, T, M, 1, TS, flags, saturated=sat_g)
    torch.cuda.synchronize(); _capi.fused_status()
    same = bool(torch.equal(want, got)) and int(sat_w) == int(sat_g)
    bad += not same
    def t(fn, n=30):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    a = t(lambda: _capi.fused_q8(x, dv, gains, want, B, A, C, 2 * C, T, M, 1, TS, flags))
    b = t(lambda: _capi.fused_packed_q8(x, packed, gains, got, B, A, C, 2 * C, T, M, 1, TS, flags))
    print(f"A={A} C={C} T={T} M={M} B={B} flags={flags}: identical {same} (saturated {int(sat_w)} / {int(sat_g)}) | q8 {a:.1f} us, packed q8 {b:.1f} us", flush=True)
sys.exit(1 if bad else 0)
