"""Developer probe: the fused kernel at the per-GPU shares of the BASELINE configurations (strong scaling: a FIXED
band cut over N GPUs, reference coeff_generator.py:53), one heap and 8 heaps per launch, default and
DCBF_FLAG_STREAMING.  Rotating buffer sets keep consecutive launches out of each other's L2 footprint.

    python tools/shares.py [--steps 40] [--batches 1,8] [name | A,C,T,M,N,xeng_id ...]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

SHARES = {
    # name: (n_ants, n_chans on this GPU, n_samples, n_beams, n_chans_total, xeng_id)
    "c3_n1": (64, 4096, 256, 64, 4096, 0),
    "c3_n2": (64, 2048, 256, 64, 4096, 1),
    "c3_n4": (64, 1024, 256, 64, 4096, 3),
    "c3_n8": (64, 512, 256, 64, 4096, 7),
    "c2_n1": (64, 1024, 256, 16, 1024, 0),
    "c4_n8": (80, 4096, 256, 32, 32768, 7),
    "c5_n8": (197, 512, 256, 256, 4096, 7),
}


def main():
    steps = 40
    batches = (1, 8)
    names = []
    args = sys.argv[1:]
    while args:
        a = args.pop(0)
        if a == "--steps":
            steps = int(args.pop(0))
        elif a == "--batches":
            batches = tuple(int(v) for v in args.pop(0).split(","))
        else:
            if "," in a:
                SHARES[a] = tuple(int(v) for v in a.split(","))
            names.append(a)
    dev = torch.device("cuda", 0)
    peak, _ = bench._peaks()
    rows = {}
    for name in names or SHARES:
        A, C, T, M, N, xid = SHARES[name]
        for B in batches:
            r = bench.measure_share(A, C, T, M, B, N, xid, steps, dev, peak)
            rows[f"{name}_b{B}"] = r
            print(f"{name:6s} B={B}: {r['ms_per_step'] * 1e3:8.1f} us {r['roofline_frac']:.3f} | streaming "
                  f"{r['streaming_ms_per_step'] * 1e3:8.1f} us {r['streaming_roofline_frac']:.3f} | graph "
                  f"{r['graph_ms_per_step'] * 1e3:8.1f} us {r['graph_roofline_frac']:.3f} | graph+streaming "
                  f"{r['graph_streaming_ms_per_step'] * 1e3:8.1f} us {r['graph_streaming_roofline_frac']:.3f} | fp16 coeff + streaming "
                  f"{r['fp16_coeff_streaming_ms_per_step'] * 1e3:8.1f} us {r['fp16_coeff_streaming_roofline_frac']:.3f} | packed "
                  f"{r['packed_ms_per_step'] * 1e3:8.1f} us {r['packed_roofline_frac'] or 0:.3f} | packed + streaming "
                  f"{r['packed_streaming_ms_per_step'] * 1e3:8.1f} us {r['packed_streaming_roofline_frac'] or 0:.3f} | host "
                  f"{r['host_us_per_launch']:6.1f} us/launch | ideal {r['algorithmic_bytes_per_launch'] / peak / 1e3:8.1f} us",
                  flush=True)
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
