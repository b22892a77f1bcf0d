"""Developer probe: same-box A/B of dcbf_fused across several builds of libdcbf.so at arbitrary shapes.

    python tools/ab_shares.py [--rounds 4] [--steps 40] [--flags 0x4] [--q8] --shape A,C,T,M,B [--shape ...] lib1.so lib2.so ...

Interleaved rounds (lib1, lib2, lib1, ...) per shape, CUDA events around `steps` launches, rotating buffer sets so
consecutive launches never find their data in L2.  Box-to-box and minute-to-minute spread (power capping) is larger than
most kernel changes, so only the interleaved numbers of one run are comparable.
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TS = 1 / 1712e6


def main():
    rounds, steps, flag_list, shapes, paths, q8 = 4, 40, [0], [], [], False
    args = sys.argv[1:]
    while args:
        a = args.pop(0)
        if a == "--rounds":
            rounds = int(args.pop(0))
        elif a == "--steps":
            steps = int(args.pop(0))
        elif a == "--flags":
            flag_list = [int(v, 0) for v in args.pop(0).split(",")]
        elif a == "--shape":
            shapes.append(tuple(int(v) for v in args.pop(0).split(",")))
        elif a == "--q8":
            q8 = True
        else:
            paths.append(a)
    dev = torch.device("cuda", 0)
    libs = []
    for path in paths:
        lib = C.CDLL(os.path.join(ROOT, path) if not os.path.isabs(path) else path)
        lib.dcbf_fused.restype = C.c_int
        lib.dcbf_fused.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_double, C.c_uint, C.c_void_p]
        lib.dcbf_fused_q8.restype = C.c_int
        lib.dcbf_fused_q8.argtypes = [C.c_void_p] * 5 + [C.c_int] * 7 + [C.c_double, C.c_void_p, C.c_uint, C.c_void_p]
        lib.dcbf_fused_status.restype = C.c_int
        lib.dcbf_fused_status.argtypes = [C.c_void_p] * 3
        libs.append((os.path.basename(path), lib))
    stream = torch.cuda.Stream()
    for (A, Cc, T, M, B) in shapes:
        alg = B * A * Cc * T * 4 + Cc * M * A * 16 + B * 2 * Cc * T * M * (2 if q8 else 8)
        sets = max(2, -(-(4 * 126_000_000) // alg))
        xs = [torch.randint(0, 256, (B, A, Cc, T, 2, 2), dtype=torch.uint8, device=dev) for _ in range(sets)]
        dvs = []
        for _ in range(sets):
            d = torch.zeros((Cc, M, A, 4), dtype=torch.float32, device=dev)
            d[..., 0] = (torch.rand((Cc, M, A), device=dev) * 32 - 16) * TS
            d[..., 2] = (torch.rand((Cc, M, A), device=dev) * 2 - 1) * 3.14159265
            dvs.append(d)
        outs = [torch.empty((B, 2, Cc, T // 16, 16, 2 * M), dtype=torch.int8 if q8 else torch.float32, device=dev) for _ in range(sets)]
        gains = torch.full((M,), 0.004, dtype=torch.float32, device=dev)

        def run(lib, n, flags):
            for i in range(n):
                j = i % sets
                if q8:
                    st = lib.dcbf_fused_q8(xs[j].data_ptr(), dvs[j].data_ptr(), gains.data_ptr(), outs[j].data_ptr(), None, B, A,
                                           Cc, Cc, T, M, 0, TS, None, flags, stream.cuda_stream)
                else:
                    st = lib.dcbf_fused(xs[j].data_ptr(), dvs[j].data_ptr(), outs[j].data_ptr(), B, A, Cc, Cc, T, M, 0, TS,
                                        flags, stream.cuda_stream)
                assert st == 0, st

        torch.cuda.synchronize()  # the inputs were generated on the default stream, the launches go to `stream`
        results = {}
        ref = None
        for name, lib in libs:  # warm-up + identity of results across builds
            run(lib, 2 * sets, 0)
            stream.synchronize()
            assert lib.dcbf_fused_status(None, None, None) == 0
            if ref is None:
                ref = outs[0].clone()
            else:
                same = torch.equal(ref, outs[0])
                print(f"  [{name}] result identical to {libs[0][0]}: {same}")
        for rnd in range(rounds):
            for flags in flag_list:
                for name, lib in (libs if rnd % 2 == 0 else libs[::-1]):  # alternate the order: clocks drift under load
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    run(lib, steps, flags)
                    e1.record(stream)
                    stream.synchronize()
                    results.setdefault((name, flags), []).append(e0.elapsed_time(e1) / steps * 1e3)
        for (name, flags), v in results.items():
            best = min(v)
            print(f"A={A} C={Cc} T={T} M={M} B={B}{' q8' if q8 else ''} flags={flags:#x} {name:24s} "
                  f"min {best:7.1f} us  all {' '.join(f'{t:7.1f}' for t in v)}  (alg {alg / best / 1e3:7.0f} GB/s)", flush=True)
        for name, lib in libs:
            assert lib.dcbf_fused_status(None, None, None) == 0
        del xs, dvs, outs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
