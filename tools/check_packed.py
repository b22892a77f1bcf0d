"""Developer probe: dcbf_fused_pack_coeffs + dcbf_fused_packed against dcbf_fused (must be bit-identical) and timings.

    python tools/check_packed.py [--time] A,C,T,M,B[,flags] ...
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402

TS = 1 / 1712e6


def main():
    args = sys.argv[1:]
    timing = "--time" in args
    shapes = [tuple(int(v, 0) for v in a.split(",")) for a in args if not a.startswith("--")]
    dev = torch.device("cuda", 0)
    bad = 0
    for sh in shapes:
        A, C, T, M, B = sh[:5]
        flags = sh[5] if len(sh) > 5 else 0
        n_total, xid = 2 * C, 1
        x = torch.randint(0, 256, (B, A, C, T, 2, 2), dtype=torch.uint8, device=dev)
        dv = torch.zeros((C, M, A, 4), dtype=torch.float32, device=dev)
        dv[..., 0] = (torch.rand((C, M, A), device=dev) * 32 - 16) * TS
        dv[..., 2] = (torch.rand((C, M, A), device=dev) * 2 - 1) * 3.14159265
        want = torch.empty((B, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
        got = torch.full_like(want, float("nan"))
        nbytes = _capi.fused_packed_bytes(A, C, M, flags)
        if not nbytes:
            print(f"A={A} C={C} T={T} M={M} B={B} flags={flags:#x}: no whole tile set (unsupported)")
            continue
        packed = torch.full((nbytes,), 0xAB, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        _capi.fused(x, dv, want, B, A, C, n_total, T, M, xid, TS, flags)
        _capi.fused_pack_coeffs(dv, packed, A, C, n_total, M, xid, TS, flags & _capi.FLAG_FP16_COEFF)
        _capi.fused_packed(x, packed, got, B, A, C, n_total, T, M, xid, TS, flags)
        torch.cuda.synchronize()
        _capi.fused_status()
        same = bool(torch.equal(want, got))
        bad += not same
        line = f"A={A} C={C} T={T} M={M} B={B} flags={flags:#x}: packed {nbytes / C / 1024:.0f} KiB/channel, identical {same}"
        if not same:
            d = (want != got) | torch.isnan(got)
            line += f" ({int(d.sum())} values differ, nan {int(torch.isnan(got).sum())}, first {d.nonzero()[0].tolist()})"
        if timing:
            def t(fn, n=30):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / n * 1e3
            us_f = t(lambda: _capi.fused(x, dv, want, B, A, C, n_total, T, M, xid, TS, flags))
            us_p = t(lambda: _capi.fused_packed(x, packed, got, B, A, C, n_total, T, M, xid, TS, flags))
            us_k = t(lambda: _capi.fused_pack_coeffs(dv, packed, A, C, n_total, M, xid, TS, flags & _capi.FLAG_FP16_COEFF))
            line += f" | fused {us_f:.1f} us, packed {us_p:.1f} us, pack {us_k:.1f} us"
        print(line, flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
