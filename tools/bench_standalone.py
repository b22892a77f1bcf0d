"""Times the three stand-alone operators (dcbf_reorder / dcbf_coeffs / dcbf_beamform) and raw PCIe copies.

    python tools/bench_standalone.py [c2|c3]

Prints one JSON object: per kernel the CUDA-event time, its algorithmic bytes (SURVEY.md section 8a sizes) and the
fraction of the measured HBM copy peak; plus pinned H2D / D2H bandwidth (the ceiling of bench.py's e2e figure).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402

CFG = {"c2": (64, 1024, 256, 16), "c3": (64, 4096, 256, 64)}


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    e[0].record()
    for i in range(n):
        fn()
        e[i + 1].record()
    torch.cuda.synchronize()
    return sorted(e[i].elapsed_time(e[i + 1]) for i in range(n))[n // 2] * 1e-3


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    A, C, T, M = CFG[name]
    B, P = 1, 2
    peak = 6550.1
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak = float(json.load(open(path))["hbm_gbs"])
    dev = torch.device("cuda", 0)
    x = torch.randint(0, 256, (B, A, C, T, 2, 2), dtype=torch.uint8, device=dev)
    re = torch.empty((B, 2, C, T // 16, 16, A, 2), dtype=torch.uint8, device=dev)
    dv = torch.rand((C, M, A, 4), dtype=torch.float32, device=dev) * 1e-8
    co = torch.empty((B, P, C, 2 * A, 2 * M), dtype=torch.float32, device=dev)
    out = torch.empty((B, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
    res = {"workload": name, "hbm_peak_GBps": peak}
    t = timed(lambda: _capi.reorder(x, re, B, A, C, T))
    by = 2 * x.numel()
    res["reorder"] = {"us": t * 1e6, "bytes": by, "GBps": by / t / 1e9, "frac_hbm": by / t / 1e9 / peak}
    t = timed(lambda: _capi.coeffs(dv, co, B, P, C, C, A, M, 0, 1 / 1712e6))
    by = dv.numel() * 4 + co.numel() * 4
    res["coeffs"] = {"us": t * 1e6, "bytes": by, "GBps": by / t / 1e9, "frac_hbm": by / t / 1e9 / peak}
    t = timed(lambda: _capi.beamform(re, co, out, B, C, T, A, M), n=5, warm=2)
    by = re.numel() + co.numel() * 4 + out.numel() * 4
    fl = B * 2 * C * T * 8 * A * M
    res["beamform"] = {"us": t * 1e6, "bytes": by, "GBps": by / t / 1e9, "frac_hbm": by / t / 1e9 / peak,
                       "tflops": fl / t / 1e12}
    t = timed(lambda: _capi.fused(x, dv, out, B, A, C, C, T, M, 0, 1 / 1712e6))
    by = _capi.fused_bytes(B, A, C, T, M)
    res["fused"] = {"us": t * 1e6, "bytes": by, "GBps": by / t / 1e9, "frac_hbm": by / t / 1e9 / peak}
    res["chain_over_fused"] = (res["reorder"]["us"] + res["coeffs"]["us"] + res["beamform"]["us"]) / res["fused"]["us"]
    # PCIe ceilings for the host-buffer path
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    t = timed(lambda: d.copy_(h, non_blocking=True), n=5, warm=1)
    res["pcie_h2d_GBps"] = n / t / 1e9
    t = timed(lambda: h.copy_(d, non_blocking=True), n=5, warm=1)
    res["pcie_d2h_GBps"] = n / t / 1e9
    h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
        s1.synchronize()
        s2.synchronize()

    import time
    both()
    t0 = time.perf_counter()
    for _ in range(3):
        both()
    t = (time.perf_counter() - t0) / 3
    res["pcie_duplex_each_GBps"] = n / t / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
