"""Kernel-only timing of the fused path at every BASELINE.json configuration on ONE GPU (C4/C5 are quoted for 8 GPUs
there; run whole here they check 64-bit indexing and the multi-N-tile path at full size).  Each result is checked on a
few sampled channels against the oracle.

    python tests/probes/bench_configs.py [c2 c3 c4 c5 ...]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402
from oracle import beamform_oracle as orc  # noqa: E402

CFG = {  # A, C, T, M
    "c1": (4, 64, 256, 4), "c2": (64, 1024, 256, 16), "c3": (64, 4096, 256, 64),
    "c4": (80, 32768, 256, 32), "c5": (197, 4096, 256, 256),
}
TS = orc.SAMPLE_PERIOD


def main():
    flags = 0
    if "--flags" in sys.argv:
        i = sys.argv.index("--flags")
        flags = int(sys.argv[i + 1], 0)
        del sys.argv[i:i + 2]
    names = sys.argv[1:] or ["c2", "c3", "c4", "c5"]
    peak = 6550.1
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak = float(json.load(open(path))["hbm_gbs"])
    dev = torch.device("cuda", 0)
    res = {}
    for name in names:
        A, C, T, M = CFG[name]
        B = 1
        g = torch.Generator(device=dev).manual_seed(5)
        x = torch.randint(0, 256, (B, A, C, T, 2, 2), dtype=torch.uint8, device=dev, generator=g)
        dv = torch.zeros((C, M, A, 4), dtype=torch.float32, device=dev)
        dv[..., 0] = (torch.rand((C, M, A), device=dev, generator=g) * 32 - 16) * TS
        dv[..., 2] = (torch.rand((C, M, A), device=dev, generator=g) * 2 - 1) * 3.14159265
        out = torch.empty((B, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
        for _ in range(2):
            _capi.fused(x, dv, out, B, A, C, C, T, M, 0, TS, flags)
        torch.cuda.synchronize()
        n = 5 if name in ("c4", "c5") else 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            _capi.fused(x, dv, out, B, A, C, C, T, M, 0, TS, flags)
        e1.record()
        torch.cuda.synchronize()
        _capi.fused_status()
        sec = e0.elapsed_time(e1) / 1e3 / n
        by = _capi.fused_bytes(B, A, C, T, M)
        fl = B * 2 * C * T * 8 * A * M
        # parity on sampled channels (first, middle, last)
        worst = 0.0
        for c0 in (0, C // 2, C - 2):
            xs = x[:, :, c0:c0 + 2].cpu().numpy()
            dvs = np.ascontiguousarray(dv[c0:c0 + 2].cpu().numpy())
            ref = orc.beamform_pipeline(xs, dvs, C, c0 // 2, TS)  # xeng geometry: 2 channels per engine
            bound = orc.beamform_abs_bound(orc.reorder(xs))[..., None]
            got = out[:, :, c0:c0 + 2].cpu().numpy()
            worst = max(worst, float(np.max(np.abs(got - ref) / bound)))
        res[name] = {"n_ants": A, "n_chans": C, "n_beams": M, "us": sec * 1e6, "algorithmic_bytes": by,
                     "GBps": by / sec / 1e9, "frac_hbm": by / sec / 1e9 / peak, "tflops_real_expanded": fl / sec / 1e12,
                     "tiling": _capi.fused_tiling(A, M, flags), "flags": flags, "max_err_over_sum_abs_x": worst,
                     "within_budget": worst <= 2.0 ** -10}
        del x, dv, out
        torch.cuda.empty_cache()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
