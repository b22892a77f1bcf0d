"""Randomised shape sweep of the C ABI against float64 references (developer aid; the committed parity tests are
tests/test_gpu_parity.py).

    python tests/probes/fuzz_gpu.py [n_cases] [seed] [--many-channels]

Every case draws (B, A, C, T, M, flags) and checks: dcbf_fused against the oracle pipeline (2^-10 sum|x| budget; the
observed error is printed), dcbf_reorder bit-exact, dcbf_beamform (tcgen05 or CUDA cores, whichever the shape takes)
against a float64 matmul, and the in-kernel watchdog status.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402
from oracle import beamform_oracle as orc  # noqa: E402

TS = orc.SAMPLE_PERIOD


def _capi_weights_log2(weights) -> int:
    """The power-of-two scale fused_ex would put on these weights (0: none, the only case the packed path takes)."""
    return 0  # the sweep's weights are in [0, 1.5]: fused_ex is called here without weights_log2


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    many_channels = "--many-channels" in sys.argv
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda", 0)
    worst_f, worst_b, bad, worst_tag = 0.0, 0.0, 0, ""
    for case in range(n_cases):
        B = int(rng.integers(1, 4))
        A = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 23, 31, 32, 33, 48, 63, 64, 65, 79, 80, 96, 100, 130, 197, 256, 300, 520]))
        M = int(rng.choice([1, 2, 3, 4, 5, 8, 15, 16, 17, 32, 33, 63, 64, 65, 100, 128, 130, 256]))
        T = 16 * int(rng.choice([1, 2, 3, 4, 7, 8, 9, 16, 17, 24]))
        while B * A * T * M > 3e6:  # keep the float64 oracle fast
            M = max(1, M // 2)
        C = int(rng.integers(1, 6)) if A * M > 4000 else int(rng.integers(1, 40))
        if many_channels:  # more work units than SMs: the persistent loops wrap, every ring changes phase many times
            A, M, T = min(A, 33), min(M, 17), min(T, 64)
            C = int(rng.integers(150, 900))
        signed = bool(rng.integers(0, 2))
        fp16 = bool(rng.integers(0, 4) == 0)
        n_total = C * int(rng.integers(1, 4))
        xid = int(rng.integers(0, n_total // C))
        x = orc.make_samples(B, A, C, T, seed=case)
        dv = orc.make_delay_vals_random(C, M, A, seed=1000 + case)
        dx, ddv = torch.from_numpy(x).to(dev), torch.from_numpy(dv).to(dev)
        out = torch.full((B, 2, C, T // 16, 16, 2 * M), float("nan"), dtype=torch.float32, device=dev)
        flags = (_capi.FLAG_SIGNED_INPUT if signed else 0) | (_capi.FLAG_FP16_COEFF if fp16 else 0)
        mode = ["plain", "plain", "tv", "weights", "q8", "tv+weights+q8"][int(rng.integers(0, 6))]
        batch_dt = list(rng.uniform(-2.0, 2.0, B)) if "tv" in mode else None
        weights = rng.uniform(0.0, 1.5, (M, A)).astype(np.float32) if "weights" in mode else None
        if batch_dt is not None:  # rates that move the phase by a few turns over the +-2 s
            dv[..., 1] = rng.uniform(-1e-10, 1e-10, dv.shape[:-1])
            dv[..., 3] = rng.uniform(-1.0, 1.0, dv.shape[:-1])
            ddv = torch.from_numpy(dv).to(dev)
        tag = f"case {case}: B{B} A{A} C{C} T{T} M{M} N{n_total} x{xid} signed={signed:d} fp16={fp16:d} {mode}"
        ref = orc.beamform_pipeline(x, dv, n_total, xid, TS, signed_input=signed, batch_dt=batch_dt, weights=weights)
        gains = out8 = sat = None
        if "q8" in mode:
            gains = (rng.uniform(0.5, 3.0, M) * 100.0 / max(np.abs(ref).max(), 1e-9)).astype(np.float32)
            out8 = torch.full(out.shape, 77, dtype=torch.int8, device=dev)
            sat = torch.zeros(1, dtype=torch.int64, device=dev)
        try:
            if mode == "plain":
                _capi.fused(dx, ddv, out, B, A, C, n_total, T, M, xid, TS, flags)
            else:
                _capi.fused_ex(dx, ddv, None if out8 is not None else out, B, A, C, n_total, T, M, xid, TS, flags,
                               batch_dt=batch_dt, weights=None if weights is None else torch.from_numpy(weights).to(dev),
                               gains=None if gains is None else torch.from_numpy(gains).to(dev), beams_q8=out8, saturated=sat)
            torch.cuda.synchronize()
            _capi.fused_status()
        except Exception as e:  # noqa: BLE001
            expected = False  # no shape in this sweep is refused any more
            print(tag, "refused as documented" if expected else f"EXCEPTION {e}", flush=True)
            bad += not expected
            continue
        budget = 2.0 ** -10 * orc.beamform_abs_bound(orc.reorder(x), signed_input=signed)[..., None]
        if weights is not None:
            budget = budget * 1.5
        if out8 is not None:  # int8 output: at most one step from the float64 requantisation, same clip count
            want, want_clipped = orc.requantise(ref, gains)
            diff = np.abs(out8.cpu().numpy().astype(np.int32) - want.astype(np.int32))
            frac = np.count_nonzero(diff) / diff.size
            ratio = 0.0 if diff.max() <= 1 and frac <= (5e-2 if fp16 else 2e-3) and abs(int(sat.item()) - want_clipped) <= 2 + 0.05 * want_clipped else float("inf")
            if ratio:
                print(f"  q8 detail: max diff {diff.max()}, fraction differing {frac:.2e}, clipped {int(sat.item())} vs {want_clipped}", flush=True)
        else:
            got = out.cpu().numpy()
            ratio = float(np.max(np.abs(got - ref) / np.maximum(budget, 1e-30))) if not np.isnan(got).any() else float("inf")
        if ratio > worst_f:
            worst_f, worst_tag = ratio, tag
        # stand-alone ops
        re = torch.empty((B, 2, C, T // 16, 16, A, 2), dtype=torch.uint8, device=dev)
        _capi.reorder(dx, re, B, A, C, T)
        ok_re = bool(np.array_equal(re.cpu().numpy(), orc.reorder(x)))
        co = torch.randn((B, 2, C, 2 * A, 2 * M), dtype=torch.float32, device=dev)
        ob = torch.full_like(out, float("nan"))
        _capi.beamform(re, co, ob, B, C, T, A, M, _capi.FLAG_SIGNED_INPUT if signed else 0)
        torch.cuda.synchronize()
        _capi.fused_status()
        xr = (re.view(torch.int8) if signed else re).double().reshape(B, 2, C, T, 2 * A)
        refb = torch.matmul(xr, co.double())
        scale = torch.matmul(xr.abs(), co.double().abs()) + 1e-30
        eb = float(((ob.reshape(B, 2, C, T, 2 * M).double() - refb).abs() / scale).max()) if not torch.isnan(ob).any() else float("inf")
        worst_b = max(worst_b, eb)
        # stand-alone coefficients (float64 evaluation; same per-heap times / weights as the fused call)
        cg = torch.full((B, 2, C, 2 * A, 2 * M), float("nan"), dtype=torch.float32, device=dev)
        _capi.coeffs(ddv, cg, B, 2, C, n_total, A, M, xid, TS, batch_dt=batch_dt,
                     weights=None if weights is None else torch.from_numpy(weights).to(dev))
        cref = orc.steering_coeffs(dv, B, 2, C, n_total, A, M, xid, TS, out_dtype=np.float64, batch_dt=batch_dt, weights=weights)
        ec = float(np.abs(cg.cpu().numpy().astype(np.float64) - cref).max())
        # host-buffer plan with a random chunking: bit-equal to the device path (plain mode only)
        ok_plan = True
        if mode == "plain":
            plan = _capi.HostPlan(B, A, C, n_total, T, M, xid, TS, flags, chunk_chans=int(rng.integers(1, C + 1)),
                                  n_slots=int(rng.integers(2, 5)))
            h_out = np.empty(out.shape, np.float32)
            plan.run(x, dv, h_out)
            plan.close()
            ok_plan = bool(np.array_equal(h_out, out.cpu().numpy()))
        # packed steering coefficients (shapes with a whole tile set): bit-equal to the per-call result, weights included
        ok_packed = True
        if mode in ("plain", "weights", "q8") and _capi.fused_packed_bytes(A, C, M, flags):
            pk = torch.full((_capi.fused_packed_bytes(A, C, M, flags),), 0xEE, dtype=torch.uint8, device=dev)
            if mode == "q8":
                dg = torch.from_numpy(gains).to(dev)
                o2 = torch.full(out.shape, 55, dtype=torch.int8, device=dev)
                sat2 = torch.zeros(1, dtype=torch.int64, device=dev)
                torch.cuda.synchronize()
                _capi.fused_pack_coeffs_q8(ddv, dg, pk, A, C, n_total, M, xid, TS, flags & _capi.FLAG_FP16_COEFF)
                _capi.fused_packed_q8(dx, pk, dg, o2, B, A, C, n_total, T, M, xid, TS, flags, saturated=sat2)
                torch.cuda.synchronize()
                ok_packed = bool(torch.equal(o2, out8)) and int(sat2.item()) == int(sat.item())
            elif weights is None or _capi_weights_log2(weights) == 0:
                o2 = torch.full_like(out, float("nan"))
                torch.cuda.synchronize()
                _capi.fused_pack_coeffs(ddv, pk, A, C, n_total, M, xid, TS, flags & _capi.FLAG_FP16_COEFF,
                                        weights=None if weights is None else torch.from_numpy(weights).to(dev))
                _capi.fused_packed(dx, pk, o2, B, A, C, n_total, T, M, xid, TS, flags)
                torch.cuda.synchronize()
                ok_packed = bool(torch.equal(o2, out))
            _capi.fused_status()
        ok = ratio <= 1.0 and ok_re and eb < 4e-6 and ec <= 2.5e-6 and ok_plan and ok_packed
        bad += not ok
        if not ok or case % 20 == 0 or A > 512:
            print(tag, f"fused err/budget {ratio:.2e} reorder {'ok' if ok_re else 'BAD'} beamform {eb:.2e} coeffs {ec:.1e} "
                       f"plan {'ok' if ok_plan else 'BAD'} packed {'ok' if ok_packed else 'BAD'} {'ok' if ok else 'FAIL'}", flush=True)
    print(f"{n_cases} cases, {bad} failed; worst fused err/budget {worst_f:.2e} ({worst_tag}), worst beamform err/sum|x||w| {worst_b:.2e}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
