"""Developer check of the tcgen05 stand-alone beamform kernel (csrc/beamform_tc.cu) against the float32 CUDA-core
kernel and a float64 torch evaluation, on a list of shapes; then its timing at C2 / C3.

    python tools/check_beamform_tc.py [--time] | --once c3
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402

SHAPES = [  # B, C, T, A, M, signed
    (1, 3, 128, 16, 16, False), (2, 3, 64, 64, 16, False), (1, 5, 256, 64, 64, False), (1, 2, 384, 80, 32, True),
    (1, 2, 256, 8, 2, False), (1, 3, 640, 72, 6, False), (1, 2, 256, 136, 130, False), (1, 300, 256, 64, 64, False),
    (1, 2, 48, 24, 70, True),
    # antenna counts whose rows are not a multiple of 16 bytes (eight-box fetch)
    (1, 3, 256, 197, 256, False), (2, 3, 64, 4, 2, False), (1, 5, 384, 79, 2, True), (1, 2, 48, 23, 4, False), (1, 300, 256, 84, 16, False),
]


CFG = {"c2": (64, 1024, 256, 16), "c3": (64, 4096, 256, 64), "c4/8": (80, 4096, 256, 32), "c2x8": (64, 8192, 256, 16),
       "c4": (80, 32768, 256, 32), "c5/8": (197, 512, 256, 256)}


def once(name):
    """three launches at a named configuration (what the ncu captures run)"""
    dev = torch.device("cuda", 0)
    A, C, T, M = CFG[name]
    re = torch.randint(0, 256, (1, 2, C, T // 16, 16, A, 2), dtype=torch.uint8, device=dev)
    co = torch.randn((1, 2, C, 2 * A, 2 * M), dtype=torch.float32, device=dev)
    out = torch.empty((1, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
    for _ in range(3):
        _capi.beamform(re, co, out, 1, C, T, A, M, 0)
    torch.cuda.synchronize()
    print("status", _capi.fused_status())


def main():
    if "--once" in sys.argv:
        return once(sys.argv[sys.argv.index("--once") + 1])
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(11)
    bad = 0
    shapes = SHAPES
    if "--shape" in sys.argv:  # --shape B C T A M signed(0/1)
        i = sys.argv.index("--shape")
        v = [int(x) for x in sys.argv[i + 1:i + 7]]
        shapes = [(v[0], v[1], v[2], v[3], v[4], bool(v[5]))]
    for (B, C, T, A, M, signed) in shapes:
        re = torch.randint(0, 256, (B, 2, C, T // 16, 16, A, 2), dtype=torch.uint8, device=dev, generator=g)
        co = torch.randn((B, 2, C, 2 * A, 2 * M), dtype=torch.float32, device=dev, generator=g)
        co *= torch.exp2(torch.randint(-20, 20, co.shape, device=dev, generator=g).float())  # wide dynamic range
        out_tc = torch.full((B, 2, C, T // 16, 16, 2 * M), float("nan"), dtype=torch.float32, device=dev)
        out_cc = torch.empty_like(out_tc)
        fl = _capi.FLAG_SIGNED_INPUT if signed else 0
        _capi.beamform(re, co, out_tc, B, C, T, A, M, fl)
        _capi.beamform(re, co, out_cc, B, C, T, A, M, fl | _capi.FLAG_DEBUG_CUDA_CORES)
        torch.cuda.synchronize()
        st = _capi.fused_status()
        x = re.view(torch.int8).double() if signed else re.double()
        x = x.reshape(B, 2, C, T, 2 * A)
        ref = torch.matmul(x, co.double())
        scale = torch.matmul(x.abs(), co.double().abs())
        got = out_tc.reshape(B, 2, C, T, 2 * M).double()
        err_tc = ((got - ref).abs() / (scale + 1e-30)).max().item()
        err_cc = ((out_cc.reshape(B, 2, C, T, 2 * M).double() - ref).abs() / (scale + 1e-30)).max().item()
        ok = err_tc < 4e-6 and not torch.isnan(out_tc).any().item()
        bad += not ok
        print(f"B{B} C{C} T{T} A{A} M{M} signed={signed}: status {st} err_tc {err_tc:.2e} err_cuda_cores {err_cc:.2e} {'ok' if ok else 'FAIL'}", flush=True)
    if "--time" in sys.argv:
        keep = []  # earlier buffers stay allocated so every repeat lands on different addresses
        for name, (A, C, T, M) in CFG.items():
            for fl, nm in ((0, "tcgen05"), (_capi.FLAG_DEBUG_CUDA_CORES, "cuda cores")):
                ts = []
                for rep in range(1 if fl else 4):
                    re = torch.randint(0, 256, (1, 2, C, T // 16, 16, A, 2), dtype=torch.uint8, device=dev)
                    co = torch.randn((1, 2, C, 2 * A, 2 * M), dtype=torch.float32, device=dev)
                    out = torch.empty((1, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
                    by = re.numel() + 4 * co.numel() + 4 * out.numel()
                    for _ in range(3):
                        _capi.beamform(re, co, out, 1, C, T, A, M, fl)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(10):
                        _capi.beamform(re, co, out, 1, C, T, A, M, fl)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) / 10 * 1e-3)
                    if by < (1 << 31):
                        keep.append(torch.empty(37 << 20, dtype=torch.uint8, device=dev))
                    del re, co, out
                t = sorted(ts)[len(ts) // 2]
                print(f"{name} {nm}: {t * 1e6:.1f} us (runs {' '.join(f'{x * 1e6:.1f}' for x in ts)}), {by / t / 1e9:.0f} GB/s = "
                      f"{by / t / 1e9 / 6550.1:.3f} of HBM copy peak", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
