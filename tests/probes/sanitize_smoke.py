"""Small run of every libdcbf entry point for compute-sanitizer (memcheck / racecheck): tiny shapes, every
code path (TMA-store and direct epilogues, ragged shapes, several N tiles, time-varying, int8 output, host plan)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402
from oracle import beamform_oracle as orc  # noqa: E402

TS = orc.SAMPLE_PERIOD
dev = torch.device("cuda", 0)
ok = True
for (B, A, C, T, M, N, xid, flags) in [(1, 4, 5, 256, 4, 64, 0, 0), (2, 5, 3, 32, 3, 256, 1, 0),
                                        (1, 64, 3, 256, 64, 4096, 0, 0), (1, 197, 2, 144, 70, 4096, 1, 0),
                                        (1, 16, 200, 64, 8, 1024, 2, _capi.FLAG_SIGNED_INPUT),
                                        (1, 33, 4, 48, 6, 128, 0, _capi.FLAG_FP16_COEFF)]:
    x = orc.make_samples(B, A, C, T, seed=A)
    dv = orc.make_delay_vals_random(C, M, A, seed=M)
    dx, ddv = torch.from_numpy(x).to(dev), torch.from_numpy(dv).to(dev)
    out = torch.zeros((B, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
    _capi.fused(dx, ddv, out, B, A, C, N, T, M, xid, TS, flags)
    _capi.fused_status()
    ref = orc.beamform_pipeline(x, dv, N, xid, TS, signed_input=bool(flags & _capi.FLAG_SIGNED_INPUT))
    bound = orc.beamform_abs_bound(orc.reorder(x), bool(flags & _capi.FLAG_SIGNED_INPUT))[..., None]
    good = bool(np.all(np.abs(out.cpu().numpy() - ref) <= 2.0 ** -10 * bound))
    ok &= good
    print(f"fused B{B} A{A} C{C} T{T} M{M} flags={flags:#x}: {'ok' if good else 'MISMATCH'}")
    # time-varying and int8 variants on the same inputs
    _capi.fused(dx, ddv, out, B, A, C, N, T, M, xid, TS, flags, batch_dt=[0.5 * b for b in range(B)])
    gains = torch.full((M,), 0.01, dtype=torch.float32, device=dev)
    out8 = torch.zeros(out.shape, dtype=torch.int8, device=dev)
    sat = torch.zeros(1, dtype=torch.int64, device=dev)
    _capi.fused_q8(dx, ddv, gains, out8, B, A, C, N, T, M, xid, TS, flags, saturated=sat)
    _capi.fused_status()
    # stand-alone operators
    re = torch.empty((B, 2, C, T // 16, 16, A, 2), dtype=torch.uint8, device=dev)
    co = torch.empty((B, 2, C, 2 * A, 2 * M), dtype=torch.float32, device=dev)
    _capi.reorder(dx, re, B, A, C, T)
    _capi.coeffs(ddv, co, B, 2, C, N, A, M, xid, TS)
    _capi.beamform(re, co, out, B, C, T, A, M, flags & _capi.FLAG_SIGNED_INPUT)
    torch.cuda.synchronize()
    ok &= bool(np.array_equal(re.cpu().numpy(), orc.reorder(x)))
plan = _capi.HostPlan(1, 8, 20, 64, 32, 4, 0, TS, chunk_chans=6, n_slots=2)
x = orc.make_samples(1, 8, 20, 32)
dv = orc.make_delay_vals_random(20, 4, 8)
host = np.zeros((1, 2, 20, 2, 16, 8), np.float32)
plan.run(x, dv, host)
plan.close()
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
