"""Developer probe: the same dcbf_fused launch repeated many times must give bit-identical beams every time.

    python tools/stress_determinism.py [--lib path.so] [--reps 400] [--flags 0x0] A,C,T,M,B [A,C,T,M,B ...]

Launches are queued back to back on one stream over rotating input / output sets (as a pipeline would), every output is
compared with the first result of its input set.  A difference would mean a race between roles of the kernel (a tile
overwritten before its last reader is done) that the parity tests, which launch once, can miss.
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpdk_dc_sand_b200 import _capi  # noqa: E402

TS = 1 / 1712e6


def main():
    args = sys.argv[1:]
    lib_path, reps, flags, shapes = None, 400, 0, []
    while args:
        a = args.pop(0)
        if a == "--lib":
            lib_path = args.pop(0)
        elif a == "--reps":
            reps = int(args.pop(0))
        elif a == "--flags":
            flags = int(args.pop(0), 0)
        else:
            shapes.append(tuple(int(v) for v in a.split(",")))
    if lib_path:
        os.environ["DCBF_LIB"] = lib_path if os.path.isabs(lib_path) else os.path.join(ROOT, lib_path)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream()
    bad_total = 0
    for (A, Cc, T, M, B) in shapes:
        sets = 3
        xs = [torch.randint(0, 256, (B, A, Cc, T, 2, 2), dtype=torch.uint8, device=dev) for _ in range(sets)]
        dvs = []
        for _ in range(sets):
            d = torch.zeros((Cc, M, A, 4), dtype=torch.float32, device=dev)
            d[..., 0] = (torch.rand((Cc, M, A), device=dev) * 32 - 16) * TS
            d[..., 2] = (torch.rand((Cc, M, A), device=dev) * 2 - 1) * 3.14159265
            dvs.append(d)
        shape = (B, 2, Cc, T // 16, 16, 2 * M)
        torch.cuda.synchronize()  # the inputs were generated on the default stream, the launches go to `stream`
        refs = []
        for j in range(sets):
            o = torch.empty(shape, dtype=torch.float32, device=dev)
            _capi.fused(xs[j], dvs[j], o, B, A, Cc, Cc, T, M, 0, TS, flags, stream)
            stream.synchronize()
            refs.append(o)
        outs = [torch.full(shape, float("nan"), dtype=torch.float32, device=dev) for _ in range(reps)]
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            for i in range(reps):
                _capi.fused(xs[i % sets], dvs[i % sets], outs[i], B, A, Cc, Cc, T, M, 0, TS, flags, stream)
        stream.synchronize()
        _capi.fused_status()
        bad = [i for i in range(reps) if not torch.equal(outs[i], refs[i % sets])]
        bad_total += len(bad)
        detail = ""
        if bad:
            i = bad[0]
            diff = (outs[i] != refs[i % sets]) | torch.isnan(outs[i])
            idx = diff.nonzero()
            detail = f"  first bad launch {i}: {int(diff.sum())} values differ, first at {idx[0].tolist()}, channels {sorted(set(idx[:, 2].tolist()))[:8]}"
        print(f"A={A} C={Cc} T={T} M={M} B={B} flags={flags:#x}: {reps} launches, {len(bad)} differ{detail}", flush=True)
        del outs, refs
        torch.cuda.empty_cache()
    sys.exit(1 if bad_total else 0)


if __name__ == "__main__":
    main()
