"""Developer probe: same-box A/B of dcbf_fused (C3) across several builds of libdcbf.so.

    python tools/ab_fused.py [--q8] [--shape A,C,T,M] lib1.so lib2.so ...      (interleaved rounds, CUDA events, 20 launches each)
"""
import ctypes as C
import sys

import torch

Q8 = "--q8" in sys.argv
if Q8:
    sys.argv.remove("--q8")
SAT = "--sat" in sys.argv  # --q8 with the saturation counter (the slower clamp)
if SAT:
    sys.argv.remove("--sat")
LONG = "--long" in sys.argv  # 3000 launches of run-in per measurement: the board is then at its power cap
if LONG:
    sys.argv.remove("--long")
A, Cc, T, M, B = 64, 4096, 256, 64, 1
if "--shape" in sys.argv:
    i = sys.argv.index("--shape")
    A, Cc, T, M = (int(v) for v in sys.argv[i + 1].split(","))
    del sys.argv[i:i + 2]
dev = torch.device("cuda", 0)
x = torch.randint(0, 256, (B, A, Cc, T, 2, 2), dtype=torch.uint8, device=dev)
dv = torch.rand((Cc, M, A, 4), dtype=torch.float32, device=dev) * 1e-8
out = torch.empty((B, 2, Cc, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
out8 = torch.empty((B, 2, Cc, T // 16, 16, 2 * M), dtype=torch.int8, device=dev)
gains = torch.full((M,), 0.01, dtype=torch.float32, device=dev)
sat = torch.zeros(1, dtype=torch.int64, device=dev)
libs = []
for path in sys.argv[1:]:
    lib = C.CDLL(path)
    lib.dcbf_fused.restype = C.c_int
    lib.dcbf_fused.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_double, C.c_uint, C.c_void_p]
    lib.dcbf_fused_q8.restype = C.c_int
    lib.dcbf_fused_q8.argtypes = [C.c_void_p] * 5 + [C.c_int] * 7 + [C.c_double, C.c_void_p, C.c_uint, C.c_void_p]
    libs.append((path, lib))


def run(lib, n):
    for _ in range(n):
        if Q8:
            st = lib.dcbf_fused_q8(x.data_ptr(), dv.data_ptr(), gains.data_ptr(), out8.data_ptr(), sat.data_ptr() if SAT else None, B, A, Cc, Cc,
                                   T, M, 0, 1 / 1712e6, None, 0, None)
        else:
            st = lib.dcbf_fused(x.data_ptr(), dv.data_ptr(), out.data_ptr(), B, A, Cc, Cc, T, M, 0, 1 / 1712e6, 0, None)
        assert st == 0, st


for path, lib in libs:
    run(lib, 5)
torch.cuda.synchronize()
for rnd in range(4):
    for path, lib in libs:
        n = 200 if LONG else 20
        if LONG:
            run(lib, 3000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(lib, n)
        e1.record()
        torch.cuda.synchronize()
        print(f"round {rnd} {path.split('/')[-1]:28s} {e0.elapsed_time(e1) / n * 1e3:7.1f} us")
