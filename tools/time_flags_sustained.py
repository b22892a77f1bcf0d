"""Developer probe: dcbf_fused across flag words (debug bits included) at the board's power cap.

    python tools/time_flags_sustained.py [--packed] A C T M flags [flags ...]

Per flag word: 0.6 s of back-to-back launches, then 200 launches between two CUDA events, SM clock and power sampled
through NVML meanwhile (the kernel follows the SM clock once the cap has lowered it).
"""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402

import pynvml  # noqa: E402

PACKED = "--packed" in sys.argv  # dcbf_fused_packed (coefficients packed once) instead of dcbf_fused
if PACKED:
    sys.argv.remove("--packed")
A, C, T, M = (int(v) for v in sys.argv[1:5])
FLAGS = [int(v, 0) for v in sys.argv[5:]] or [0]
dev = torch.device("cuda", 0)
x = torch.randint(0, 256, (1, A, C, T, 2, 2), dtype=torch.uint8, device=dev)
dv = torch.rand((C, M, A, 4), dtype=torch.float32, device=dev) * 1e-8
out = torch.empty((1, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
byts = _capi.load().dcbf_fused_bytes(1, A, C, T, M)
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


packed = {}


def launch(f, n):
    if PACKED:
        key = f & _capi.FLAG_FP16_COEFF
        if key not in packed:
            packed[key] = torch.empty(_capi.fused_packed_bytes(A, C, M, key), dtype=torch.uint8, device=dev)
            _capi.fused_pack_coeffs(dv, packed[key], A, C, C, M, 0, 1 / 1712e6, key)
        for _ in range(n):
            _capi.fused_packed(x, packed[key], out, 1, A, C, C, T, M, 0, 1 / 1712e6, f)
        return
    for _ in range(n):
        _capi.fused(x, dv, out, 1, A, C, C, T, M, 0, 1 / 1712e6, f)


for f in FLAGS:
    launch(f, 5)
torch.cuda.synchronize()
for rnd in range(2):
    for f in FLAGS:
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 0.6:
            launch(f, 100)
            torch.cuda.synchronize()
        samples, stop = [], threading.Event()

        def sample():
            while not stop.is_set():
                samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
                time.sleep(0.002)

        th = threading.Thread(target=sample)
        th.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch(f, 200)
        e1.record()
        torch.cuda.synchronize()
        stop.set()
        th.join()
        us = e0.elapsed_time(e1) / 200 * 1e3
        mhz = sorted(s[0] for s in samples)[len(samples) // 2]
        watt = sorted(s[1] for s in samples)[len(samples) // 2]
        print(f"round {rnd} flags {f:#9x}: {us:8.1f} us  {byts / us / 1e3 / 6550.1:.3f} of 6550 GB/s  {mhz} MHz {watt:.0f} W  ({us * mhz / 1e3:.0f} kcycles)", flush=True)
        time.sleep(0.5)
_capi.fused_status()
