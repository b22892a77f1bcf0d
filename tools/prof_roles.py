"""Developer probe: per-warp-role blocked time inside the fused kernel (dcbf_debug_set_profile_buffer)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpdk_dc_sand_b200 import _capi  # noqa: E402

ROLES = {1: "producer", 2: "mma", 3: "epilogue", 4: "convert", 5: "coeff"}
SLOTS = {1: ["raw_empty", "-", "-"], 2: ["bop_full", "acc_empty", "aop_full"], 3: ["acc_full", "store_wait", "tmem|sts+fence"],
         4: ["raw_full", "aop_empty", "work|fence"], 5: ["bop_empty", "data_wait", "first_set@"]}


def main():
    A, C, T, M, B = (int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (64, 4096, 256, 64, 1)))
    flags = int(sys.argv[6], 0) if len(sys.argv) > 6 else 0
    q8 = len(sys.argv) > 7 and sys.argv[7] == "q8"
    dev = torch.device("cuda", 0)
    x = torch.randint(0, 256, (B, A, C, T, 2, 2), dtype=torch.uint8, device=dev)
    dv = torch.rand((C, M, A, 4), dtype=torch.float32, device=dev) * 1e-8
    out = torch.empty((B, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    prof = torch.zeros(n_sm * 24, dtype=torch.int64, device=dev)
    lib = _capi.load()
    gains = torch.full((M,), 0.004, dtype=torch.float32, device=dev)
    out8 = torch.empty(out.shape, dtype=torch.int8, device=dev) if q8 else None

    def launch():
        if q8:
            _capi.fused_q8(x, dv, gains, out8, B, A, C, C, T, M, 0, 1 / 1712e6, flags)
        else:
            _capi.fused(x, dv, out, B, A, C, C, T, M, 0, 1 / 1712e6, flags)

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    lib.dcbf_debug_set_profile_buffer(prof.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    torch.cuda.synchronize()
    lib.dcbf_debug_set_profile_buffer(None)
    _capi.fused_status()
    raw = prof.cpu().numpy().reshape(n_sm, 6, 4)
    ep2 = raw[:, 3, 2].copy()
    raw[:, 3, 2] = (ep2 & 0xffffffff) + (ep2 >> 32)
    cv2 = raw[:, 4, 2].copy()
    raw[:, 4, 2] = (cv2 & 0xffffffff) + (cv2 >> 32)
    p = raw.astype(np.float64) / 1e3  # us
    print(f"  convert detail: lds+alu+sts {(cv2 & 0xffffffff).mean() / 1e3:.1f} us, fence+arrive {(cv2 >> 32).mean() / 1e3:.1f} us")
    print(f"  epilogue detail: tmem wait {(ep2 & 0xffffffff).mean() / 1e3:.1f} us, sts+fence {(ep2 >> 32).mean() / 1e3:.1f} us")
    print(f"A={A} C={C} T={T} M={M} B={B} flags={flags:#x}{' q8' if q8 else ''}: kernel {e0.elapsed_time(e1)*1e3:.1f} us; "
          f"per-role blocked time, mean over {n_sm} CTAs (us)")
    t_in, t_role, t_out = p[:, 0, 0], p[:, 0, 1], p[:, 0, 2]
    base = t_in.min()
    print(f"  CTA entry skew {t_in.max() - base:6.1f} us | prologue (entry->roles) mean {(t_role - t_in).mean():5.1f} max "
          f"{(t_role - t_in).max():5.1f} | last role start {t_role.max() - base:6.1f} | first exit {t_out.min() - base:6.1f} "
          f"last exit {t_out.max() - base:6.1f} us after first entry")
    for r, name in ROLES.items():
        span = p[:, r, 3]
        parts = ", ".join(f"{SLOTS[r][k]}={p[:, r, k].mean():7.1f}" for k in range(3) if SLOTS[r][k] != "-")
        busy = span - (p[:, r, :2].sum(axis=1) if r == 5 else p[:, r, :3].sum(axis=1))
        print(f"  {name:9s} span {span.mean():7.1f} (max {span.max():7.1f})  busy {busy.mean():7.1f}  blocked: {parts}")


if __name__ == "__main__":
    main()
