"""Host-link probe: pinned host <-> device copy rates with 1, 2, 4 and 8 GPUs of one box busy at the same time.

    python tools/link_probe.py [--mib 256] [--repeats 5] > profiles/r02_link_probe_8gpu.json

For every GPU count g and every direction (H2D, D2H, both at once) all g GPUs copy `mib` MiB buffers concurrently
(one stream per GPU and direction, CUDA events per GPU); per-GPU and aggregate GB/s are reported.  This names the limiter
of the end-to-end (`e2e`) figure of bench.py at N > 1: if the aggregate rate stops growing with g, the host fabric
(root complex / memory / IOMMU of the VM) is what the ranks share.  (Method after the reference's
utilities/pcie_bandwidth_tests/cudaPcieRateTest.cpp:63-123 -- concurrent cudaMemcpyAsync in both directions timed with
events -- not its code.)
"""
import json
import sys

import torch


def probe(devs, mib, repeats, h2d, d2h):
    n = mib * 2**20
    bufs = []
    for d in devs:
        with torch.cuda.device(d):
            bufs.append({"dev_in": torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}"),
                         "dev_out": torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}"),
                         "host_in": torch.empty(n, dtype=torch.uint8, pin_memory=True),
                         "host_out": torch.empty(n, dtype=torch.uint8, pin_memory=True),
                         "s_up": torch.cuda.Stream(device=d), "s_down": torch.cuda.Stream(device=d)})
    best = None
    for rep in range(repeats + 1):
        evs = []
        for d, b in zip(devs, bufs):
            with torch.cuda.device(d):
                e = {}
                if h2d:
                    e["u0"], e["u1"] = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    with torch.cuda.stream(b["s_up"]):
                        e["u0"].record()
                        b["dev_in"].copy_(b["host_in"], non_blocking=True)
                        e["u1"].record()
                if d2h:
                    e["d0"], e["d1"] = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    with torch.cuda.stream(b["s_down"]):
                        e["d0"].record()
                        b["host_out"].copy_(b["dev_out"], non_blocking=True)
                        e["d1"].record()
                evs.append(e)
        for d in devs:
            torch.cuda.synchronize(d)
        if rep == 0:
            continue  # warm-up
        row = {"h2d_GBps_per_gpu": [n / (e["u0"].elapsed_time(e["u1"]) / 1e3) / 1e9 for e in evs] if h2d else None,
               "d2h_GBps_per_gpu": [n / (e["d0"].elapsed_time(e["d1"]) / 1e3) / 1e9 for e in evs] if d2h else None}
        total = sum(row["h2d_GBps_per_gpu"] or []) + sum(row["d2h_GBps_per_gpu"] or [])
        if best is None or total > best[0]:
            best = (total, row)
    row = best[1]
    row["aggregate_GBps"] = best[0]
    return row


def main():
    mib = int(sys.argv[sys.argv.index("--mib") + 1]) if "--mib" in sys.argv else 256
    repeats = int(sys.argv[sys.argv.index("--repeats") + 1]) if "--repeats" in sys.argv else 5
    n_dev = torch.cuda.device_count()
    out = {"what": __doc__.split("\n\n")[0], "mib_per_copy": mib, "gpus_visible": n_dev, "rows": []}
    for g in (1, 2, 4, 8):
        if g > n_dev:
            break
        devs = list(range(g))
        for name, (h2d, d2h) in {"h2d": (True, False), "d2h": (False, True), "duplex": (True, True)}.items():
            r = probe(devs, mib, repeats, h2d, d2h)
            r = {"gpus": g, "direction": name, **r}
            for k in ("h2d_GBps_per_gpu", "d2h_GBps_per_gpu"):
                if r[k] is not None:
                    r[k] = [round(v, 2) for v in r[k]]
            r["aggregate_GBps"] = round(r["aggregate_GBps"], 2)
            out["rows"].append(r)
            sys.stderr.write(f"{g} GPUs {name:6s}: aggregate {r['aggregate_GBps']:7.1f} GB/s  "
                             f"h2d {r['h2d_GBps_per_gpu']} d2h {r['d2h_GBps_per_gpu']}\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
