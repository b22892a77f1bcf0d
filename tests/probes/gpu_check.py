"""Developer probe (run under gpurun): every libdcbf entry point vs the oracle on a list of shapes, printing
max errors instead of asserting, so that one GPU call yields a full picture.  Not part of the product."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402
from oracle import beamform_oracle as orc  # noqa: E402

TS = orc.SAMPLE_PERIOD
dev = torch.device("cuda", 0)


def run_case(B, A, C, T, M, N, xid, flags=0, uniform=False, tag=""):
    x = orc.make_samples(B, A, C, T, seed=2021 + A + M)
    dv = orc.make_delay_vals_uniform(C, M, A) if uniform else orc.make_delay_vals_random(C, M, A, seed=7 + A)
    dx, ddv = torch.from_numpy(x).to(dev), torch.from_numpy(dv).to(dev)
    out = torch.full((B, 2, C, T // 16, 16, 2 * M), float("nan"), dtype=torch.float32, device=dev)
    t0 = time.time()
    _capi.fused(dx, ddv, out, B, A, C, N, T, M, xid, TS, flags)
    try:
        _capi.fused_status()
    except Exception as exc:  # watchdog
        print(f"[fused {tag}] B{B} A{A} C{C} T{T} M{M}: FAILED {exc}")
        return False
    dt = time.time() - t0
    got = out.cpu().numpy().astype(np.float64)
    signed = bool(flags & _capi.FLAG_SIGNED_INPUT)
    ref = orc.beamform_pipeline(x, dv, N, xid, TS, signed_input=signed)
    bound = orc.beamform_abs_bound(orc.reorder(x), signed_input=signed)[..., None]
    err = np.abs(got - ref)
    n_nan = int(np.isnan(got).sum())
    rel = float(np.nanmax(err / np.maximum(bound, 1e-30)))
    ok = n_nan == 0 and rel <= 2.0 ** -10
    print(f"[fused {tag}] B{B} A{A} C{C} T{T} M{M} N{N} x{xid} flags={flags:#x}: max|err|={np.nanmax(err):.3e} "
          f"err/sum|x|={rel:.3e} (budget {2.0**-10:.3e}) nan={n_nan} tiling={_capi.fused_tiling(A, M, flags)} "
          f"{'OK' if ok else 'BAD'} ({dt*1e3:.1f} ms)")
    if not ok:
        bad = np.argwhere(~(err <= bound * 2.0 ** -10))
        print("   first bad idx:", bad[:5].tolist(), "got", got[tuple(bad[0])], "ref", ref[tuple(bad[0])])
        # which pols / rows / cols are bad?
        badmask = ~(err <= bound * 2.0 ** -10)
        print("   bad by pol:", badmask.sum(axis=(0, 2, 3, 4, 5)).tolist(), "by col:",
              badmask.sum(axis=(0, 1, 2, 3, 4)).tolist()[:16], "by t16:", badmask.sum(axis=(0, 1, 2, 3, 5)).tolist())
    return ok


def run_standalone(B, A, C, T, M, N, xid):
    x = orc.make_samples(B, A, C, T, seed=1)
    dv = orc.make_delay_vals_random(C, M, A, seed=2)
    dx, ddv = torch.from_numpy(x).to(dev), torch.from_numpy(dv).to(dev)
    re = torch.zeros((B, 2, C, T // 16, 16, A, 2), dtype=torch.uint8, device=dev)
    _capi.reorder(dx, re, B, A, C, T)
    ok_r = np.array_equal(re.cpu().numpy(), orc.reorder(x))
    co = torch.zeros((B, 2, C, 2 * A, 2 * M), dtype=torch.float32, device=dev)
    _capi.coeffs(ddv, co, B, 2, C, N, A, M, xid, TS)
    ref_co = orc.steering_coeffs(dv, B, 2, C, N, A, M, xid, TS)
    dco = np.abs(co.cpu().numpy().astype(np.float64) - ref_co.astype(np.float64)).max()
    n_neq = int((co.cpu().numpy() != ref_co).sum())
    out = torch.zeros((B, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
    _capi.beamform(re, co, out, B, C, T, A, M)
    ref = orc.beamform(orc.reorder(x), ref_co)
    e = np.abs(out.cpu().numpy() - ref)
    print(f"[standalone] B{B} A{A} C{C} T{T} M{M}: reorder exact={ok_r} coeff max diff={dco:.2e} (!= count {n_neq}) "
          f"beamform max err={e.max():.3e} rel={np.max(e / np.maximum(np.abs(ref), 1)):.2e}")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "libdcbf", _capi.lib_path())
    run_standalone(2, 5, 3, 32, 3, 256, 1)
    run_standalone(1, 64, 8, 256, 16, 1024, 0)
    ok = True
    cases = [
        # B, A, C, T, M, N, xid
        (1, 4, 8, 256, 4, 64, 0),
        (1, 64, 4, 256, 64, 4096, 0),
        (1, 64, 300, 256, 16, 1024, 0),
        (2, 5, 3, 32, 3, 256, 1),
        (1, 80, 5, 256, 32, 32768, 3),
        (3, 23, 7, 48, 2, 1024, 0),
        (1, 197, 2, 256, 256, 4096, 1),
    ]
    for c in cases:
        ok &= run_case(*c, tag="16x256b")
    ok &= run_case(1, 64, 4, 256, 64, 4096, 0, flags=_capi.FLAG_DEBUG_DIRECT_EPILOGUE, tag="direct")
    ok &= run_case(1, 64, 4, 256, 64, 4096, 0, flags=_capi.FLAG_FP16_COEFF, tag="fp16")
    ok &= run_case(1, 64, 4, 256, 64, 4096, 0, flags=_capi.FLAG_SIGNED_INPUT, tag="signed")
    ok &= run_case(1, 4, 8, 256, 4, 64, 0, uniform=True, tag="uniform")
    print("ALL OK" if ok else "SOME BAD")
