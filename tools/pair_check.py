"""Developer probe: the CTA-pair (cta_group::2) K-streamed kernel against the single-CTA one (DCBF_FLAG_DEBUG_NO_PAIR)
on a few shapes -- largest difference relative to sum|x|, and timings at the 8-GPU share of C5.

    python tools/pair_check.py [--time]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402

SHAPES = [  # B, A, C, T, M
    (1, 197, 8, 256, 256), (1, 100, 5, 384, 100), (2, 197, 3, 256, 130), (1, 520, 3, 160, 70), (1, 64, 9, 256, 200),
    (3, 33, 4, 640, 97), (1, 197, 4, 256, 125), (2, 61, 3, 384, 250), (1, 36, 2, 512, 128),
]


def run(shape, flags, reps=1, time_it=False):
    b, a, c, t, m = shape
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randint(0, 256, (b, a, c, t, 2, 2), dtype=torch.uint8, device="cuda", generator=g)
    dv = torch.rand((c, m, a, 4), dtype=torch.float32, device="cuda", generator=g) * 2e-7
    dv[..., 2] = (torch.rand((c, m, a), device="cuda", generator=g) - 0.5) * 6.0
    out = torch.full((b, 2, c, t // 16, 16, 2 * m), float("nan"), dtype=torch.float32, device="cuda")
    for _ in range(reps):
        _capi.fused(x, dv, out, b, a, c, 4096, t, m, 1, 1 / 1712e6, flags)
    torch.cuda.synchronize()
    _capi.fused_status()
    ms = None
    if time_it:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _capi.fused(x, dv, out, b, a, c, 4096, t, m, 1, 1 / 1712e6, flags)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
    bound = x.to(torch.float32).abs().sum(dim=(1, 5)).amax()  # crude: max over everything of sum|x| over antennas
    return out, float(bound), ms


def main():
    for shape in SHAPES:
        o_pair, bound, _ = run(shape, 0)
        o_one, _, _ = run(shape, _capi.FLAG_DEBUG_NO_PAIR)
        bad = int(torch.isnan(o_pair).sum())
        d = float((o_pair - o_one).abs().max()) if not bad else float("nan")
        print(f"{shape}: max|pair - single| = {d:.3e} ({d / bound:.2e} of sum|x|), NaN left: {bad}", flush=True)
    if "--time" in sys.argv:
        shape = (1, 197, 512, 256, 256)
        for name, flags in (("pair", 0), ("single", _capi.FLAG_DEBUG_NO_PAIR)):
            _, _, ms = run(shape, flags, reps=3, time_it=True)
            byts = _capi.load().dcbf_fused_bytes(*[int(v) for v in shape])
            print(f"C5 share {name}: {ms * 1e3:.1f} us  {byts / ms / 1e6 / 6550.1:.3f} of 6550 GB/s", flush=True)


if __name__ == "__main__":
    main()
