#!/usr/bin/env python
"""Benchmark of the fused tied-array beamforming hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of reorder + steering coefficients + beamform (one `dcbf_fused` launch) over one heap
batch of synthetic input resident in HBM.  Workload (default c3 = the configuration the BASELINE target is
quoted on): MeerKAT 4k mode, 64 antennas x 2 pols x 4096 channels x 256 samples, 64 beams, per GPU.
With N GPUs every rank processes its own 4096-channel stream (rank == xeng_id, n_channels = 4096*N):
frequency-channel sharding, no data-path collective -> "scaling": "weak".

One JSON line is printed by rank 0:
  value          whole-job input GB/s (all ranks' voltage bytes / max-over-ranks device time), inputs in HBM
  roofline       algorithmic bytes (in + delay_vals + out) per launch / mean launch time,
                 against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  e2e            the same metric through the host-buffer C-ABI call (dcbf_host_plan_run): pinned host arrays,
                 H2D + kernel + D2H inside the timed region
  cpu_baseline   the oracle's vectorised numpy port of the same path on this box's host cores (bounded sample)
`--impl reference` times that CPU port alone (the reference's own implementation is numba/python and cannot
travel to the GPU box) and prints the same line shape with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, n_ants, n_chans per GPU, n_samples, n_beams, n_batches)
    "c1": ("4 antennas x 2 pols x 64 channels x 256 samples, 4 beams (parity case)", 4, 64, 256, 4, 1),
    "c2": ("MeerKAT 64 antennas x 2 pols x 1024 channels x 256 samples/heap, 16 beams", 64, 1024, 256, 16, 1),
    "c3": ("MeerKAT 4k mode: 64 antennas x 2 pols x 4096 channels x 256 samples, 64 beams", 64, 4096, 256, 64, 1),
    # same geometry as c2 with 8 heaps per launch (n_batches = 8; the reference's own tests batch 3 heaps)
    "c2_b8": ("MeerKAT 64 antennas x 2 pols x 1024 channels x 256 samples/heap, 16 beams, 8 heaps per launch", 64, 1024, 256, 16, 8),
}
SAMPLE_PERIOD = 1 / 1712e6
METRIC = "fused reorder+coeff+beamform input throughput"
UNIT = "GB/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic(workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "fused_traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get(workload)
    return None


def _verbatim_cpu():
    """The reference's own CPU functions timed once in the authoring container (tools/time_verbatim_reference.py): they
    need /root/reference, which does not exist on the GPU box, so the figure is context beside the live port timing."""
    path = os.path.join(ROOT, "profiles", "r02_cpu_verbatim_reference.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        d = json.load(fh)
    return {"source": "profiles/r02_cpu_verbatim_reference.json (authoring container, not this run)", "cores": d["cores"],
            "what": d["what"], "unit": UNIT,
            "cases": [{"case": c["case"], "value": c["input_GBps"], "total_s": c["total_s"]} for c in d["cases"]]}


def _ncu_kernel_us(workload: str):
    """gpu__time_duration of one launch in the committed ncu capture (cold, serialised: context only)."""
    path = os.path.join(ROOT, "profiles", "fused_traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get(f"{workload}_detail", {}).get("gpu_time_us")
    return None


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while a timed region runs.

    NVML from a thread (one sample per millisecond: the timed region lasts only a few milliseconds), with the
    recipe's `nvidia-smi -lms` query as the fall-back when pynvml is unavailable."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index = index
        self.rows = []  # (time, sm_mhz, max_mhz, power_w, set of reasons)
        self.proc = None
        self.thread = None
        self._stop = False
        self.source = None

    def _nvml_loop(self, nv, handle) -> None:
        reasons = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                   "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                   "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                   "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        mx = nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                pw = nv.nvmlDeviceGetPowerUsage(handle) / 1e3
                self.rows.append((time.time(), float(sm), float(mx), pw, {k for k, v in reasons.items() if mask & v}))
            except Exception:
                pass
            time.sleep(0.001)

    def _smi_loop(self) -> None:
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
                rs = {n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")}
                self.rows.append((time.time(), float(f[0]), float(f[1]), float(f[2]), rs))
            except ValueError:
                continue

    def start(self) -> None:
        try:
            import pynvml as nv

            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.source = "nvml, 1 ms period"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 50"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def window(self, t_begin: float, t_end: float) -> dict:
        """Summary of the samples taken inside [t_begin, t_end] (the nearest ones if the window caught none)."""
        rows = [r for r in self.rows if t_begin <= r[0] <= t_end]
        inside = len(rows)
        if not rows and self.rows:
            mid = 0.5 * (t_begin + t_end)
            rows = sorted(self.rows, key=lambda r: abs(r[0] - mid))[:2]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples available"], "samples": 0}
        reasons = set().union(*[r[4] for r in rows])
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": max(r[2] for r in rows),
                "power_w": round(statistics.median(r[3] for r in rows), 1), "reasons": sorted(reasons),
                "samples": inside, "source": self.source}

    def stop(self) -> None:
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=1)


# --------------------------------------------------------------------------------------------------------
# CPU arm (oracle port) -- the only place bench.py touches oracle/
# --------------------------------------------------------------------------------------------------------
def cpu_port_rate(n_ants, n_beams, n_samples, n_chans_total, sample_chans, repeats, budget_s=25.0):
    """Input GB/s of the oracle's numpy port on `sample_chans` channels of the workload, on ALL host cores.

    Channels are independent, so the sample is cut into one block of channels per core and the oracle function is
    run on the blocks from a thread pool (numpy releases the GIL inside its kernels; BLAS is pinned to one thread
    per worker so the cores are not oversubscribed).  Whichever of {one call with threaded BLAS, one block per
    core} is faster on this box is the figure reported."""
    import concurrent.futures as cf

    import numpy as np

    from oracle import beamform_oracle as orc

    x = orc.make_samples(1, n_ants, sample_chans, n_samples, seed=2021)
    dv = orc.make_delay_vals_random(sample_chans, n_beams, n_ants, seed=2022)
    orc.beamform_pipeline_fast(x[:, :, :4], dv[:4], n_chans_total, 0, SAMPLE_PERIOD)  # warm-up (BLAS init)
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    n_blocks = max(1, min(cores, sample_chans))
    edges = [sample_chans * i // n_blocks for i in range(n_blocks + 1)]
    blocks = [(np.ascontiguousarray(x[:, :, lo:hi]), np.ascontiguousarray(dv[lo:hi]), lo)
              for lo, hi in zip(edges[:-1], edges[1:]) if hi > lo]

    def one_block(args):
        xb, dvb, lo = args
        # xeng geometry: this block is engine lo / len of a stream of equally sized engines -> same channel phases
        return orc.beamform_pipeline_fast(xb, dvb, n_chans_total, 0, SAMPLE_PERIOD).shape

    def run_single():
        return orc.beamform_pipeline_fast(x, dv, n_chans_total, 0, SAMPLE_PERIOD).shape

    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None

    def run_blocks(pool):
        if threadpool_limits is None:
            return list(pool.map(one_block, blocks))
        with threadpool_limits(limits=1):
            return list(pool.map(one_block, blocks))

    results = {}
    with cf.ThreadPoolExecutor(max_workers=len(blocks)) as pool:
        for name, fn in (("single call, threaded BLAS", run_single), (f"{len(blocks)} channel blocks on a thread pool", lambda: run_blocks(pool))):
            fn()
            times = []
            t_start = time.perf_counter()
            for _ in range(repeats):
                t0 = time.perf_counter()
                fn()
                times.append(time.perf_counter() - t0)
                if time.perf_counter() - t_start > budget_s / 2:
                    break
            results[name] = times
    how = min(results, key=lambda k: statistics.median(results[k]))
    times = results[how]
    best = statistics.median(times)
    cpu_port_rate.how = how
    return x.nbytes / best / 1e9, int(cores), times


def bench_config(workload, wl, world):
    """The `config` object of the JSON line: the same keys and values in both arms (ours / --impl reference)."""
    desc, A, C, T, M, B = wl
    alg = B * A * C * T * 4 + C * M * A * 16 + B * 2 * C * T * M * 8
    return {"workload": f"{workload}: {desc} per GPU", "n_ants": A, "n_chans_per_gpu": C, "n_chans_total": C * world,
            "n_samples": T, "n_beams": M, "n_batches": B,
            "parallelism": f"channel-sharded x{world} (rank == xeng_id), no collective",
            "l2": f"working set {alg / 2**20:.0f} MiB per step > 126 MB L2 (inputs larger than L2)"}


def run_reference(args, wl) -> None:
    desc, A, C, T, M, B = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = min(C, 256)
    t0 = time.perf_counter()
    # W warm-up + K timed steps, each step = one pass over the bounded sample
    rate, threads, times = cpu_port_rate(A, M, T, C * args.gpus, sample, repeats=args.warmup + args.steps, budget_s=120.0)
    timed = times[args.warmup:] or times
    sec = statistics.mean(timed)
    in_bytes = B * A * sample * T * 4
    value = in_bytes / sec / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(timed), "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.workload, wl, args.gpus), "sample": f"{sample} of {C} channels per step",
        "beam_gsamples_per_s": B * 2 * sample * T * M / sec / 1e9,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} of {C} channels ({in_bytes} input bytes) per step, numpy float32 "
                                   f"transpose + float64 phase + BLAS matmul, {getattr(cpu_port_rate, 'how', '')}; host has "
                                   f"{os.cpu_count()} logical cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------
def measure_share(A, C, T, M, B, n_total, xeng_id, steps, dev, peak, desc=None):
    """Kernel-only figure of one (per-GPU share of a) configuration on this GPU: inputs resident in HBM, `sets`
    input/output sets rotated so that consecutive launches never find their data in the 126 MB L2.  `C` channels of
    a band of `n_total`, this GPU being X-engine `xeng_id` (reference: coeff_generator.py:53)."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    alg = _capi.fused_bytes(B, A, C, T, M)
    sets = max(2, -(-(4 * 126_000_000) // alg))  # >= 4 x L2 in flight
    gen = torch.Generator(device=dev)
    gen.manual_seed(77 + xeng_id)
    xs = [torch.randint(0, 256, (B, A, C, T, 2, 2), dtype=torch.uint8, device=dev, generator=gen) for _ in range(sets)]
    dvs = []
    for _ in range(sets):
        d = torch.zeros((C, M, A, 4), dtype=torch.float32, device=dev)
        d[..., 0] = (torch.rand((C, M, A), device=dev, generator=gen) * 32 - 16) * SAMPLE_PERIOD
        d[..., 2] = (torch.rand((C, M, A), device=dev, generator=gen) * 2 - 1) * 3.14159265
        dvs.append(d)
    outs = [torch.empty((B, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev) for _ in range(sets)]
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()  # the inputs were generated on the default stream, the launches go to `stream`

    def enqueue(n, flags, st):
        for i in range(n):
            j = i % sets
            _capi.fused(xs[j], dvs[j], outs[j], B, A, C, n_total, T, M, xeng_id, SAMPLE_PERIOD, flags, st)

    # packed steering coefficients (dcbf_fused_pack_coeffs once per delay model, dcbf_fused_packed per heap); shapes that
    # keep no whole tile set (C5) have no such path
    packs = None
    if _capi.fused_packed_bytes(A, C, M, 0):
        packs = [torch.empty(_capi.fused_packed_bytes(A, C, M, 0), dtype=torch.uint8, device=dev) for _ in range(sets)]
        for j in range(sets):
            _capi.fused_pack_coeffs(dvs[j], packs[j], A, C, n_total, M, xeng_id, SAMPLE_PERIOD, 0, stream)
        stream.synchronize()

    def enqueue_packed(n, flags, st):
        for i in range(n):
            j = i % sets
            _capi.fused_packed(xs[j], packs[j], outs[j], B, A, C, n_total, T, M, xeng_id, SAMPLE_PERIOD, flags, st)

    def run(n, flags, enqueue=enqueue):
        """(device seconds per launch, host seconds spent issuing one launch) of n launches from the host."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            t0 = time.perf_counter()
            enqueue(n, flags, stream)
            t_host = (time.perf_counter() - t0) / n
            e1.record(stream)
        stream.synchronize()
        return e0.elapsed_time(e1) / 1e3 / n, t_host

    def run_graph(n, flags):
        """Device seconds per launch of the same n launches captured once into a CUDA graph and replayed: what the
        GPU needs when the host's per-launch cost (argument checks, two tensor-map encodes, the launch itself) is
        off the critical path -- a launch of a small share is shorter than that cost."""
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            enqueue(n, flags, stream)
        graph.replay()
        torch.cuda.synchronize()
        time.sleep(0.25)  # (the same idle as before the live bursts: two replays back to back reach the power cap)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        del graph
        return e0.elapsed_time(e1) / 1e3 / n

    # a short idle before every timed burst: after a few tenths of a second of back-to-back launches the board sits at
    # its power cap (the `sustained` figure of the headline shows by how much), and these figures are compared with the
    # burst copy peak like the headline is
    run(2 * sets, 0)
    time.sleep(0.25)
    run(sets, 0)
    sec, host_sec = run(steps, 0)
    time.sleep(0.25)
    run(sets, _capi.FLAG_STREAMING)
    sec_stream, _ = run(steps, _capi.FLAG_STREAMING)
    time.sleep(0.25)
    sec_graph = run_graph(steps, 0)
    time.sleep(0.25)
    sec_graph_stream = run_graph(steps, _capi.FLAG_STREAMING)
    time.sleep(0.25)
    # the optional single fp16 rounding of the coefficients (one MMA pass, inside the 2^-10 sum|x| budget) with streaming
    run(sets, _capi.FLAG_STREAMING | _capi.FLAG_FP16_COEFF)
    sec_h_stream, _ = run(steps, _capi.FLAG_STREAMING | _capi.FLAG_FP16_COEFF)
    sec_p = sec_p_stream = 0.0
    if packs is not None:
        time.sleep(0.25)
        run(sets, 0, enqueue_packed)
        sec_p, _ = run(steps, 0, enqueue_packed)
        time.sleep(0.25)
        run(sets, _capi.FLAG_STREAMING, enqueue_packed)
        sec_p_stream, _ = run(steps, _capi.FLAG_STREAMING, enqueue_packed)
    _capi.fused_status()
    out = {"ms_per_step": sec * 1e3, "input_GBps": xs[0].numel() / sec / 1e9,
           "beam_gsamples_per_s": B * 2 * C * T * M / sec / 1e9, "algorithmic_bytes_per_launch": alg,
           "roofline_frac": alg / sec / 1e9 / peak, "streaming_ms_per_step": sec_stream * 1e3,
           "streaming_roofline_frac": alg / sec_stream / 1e9 / peak,
           "host_us_per_launch": host_sec * 1e6,
           "graph_ms_per_step": sec_graph * 1e3, "graph_roofline_frac": alg / sec_graph / 1e9 / peak,
           "graph_streaming_ms_per_step": sec_graph_stream * 1e3,
           "graph_streaming_roofline_frac": alg / sec_graph_stream / 1e9 / peak,
           "fp16_coeff_streaming_ms_per_step": sec_h_stream * 1e3,
           "fp16_coeff_streaming_roofline_frac": alg / sec_h_stream / 1e9 / peak,
           "packed_ms_per_step": sec_p * 1e3, "packed_roofline_frac": alg / sec_p / 1e9 / peak if sec_p else None,
           "packed_streaming_ms_per_step": sec_p_stream * 1e3,
           "packed_streaming_roofline_frac": alg / sec_p_stream / 1e9 / peak if sec_p_stream else None,
           "l2": f"{sets} rotating input/output sets",
           "geometry": {"n_ants": A, "n_chans_per_gpu": C, "n_chans_total": n_total, "n_samples": T, "n_beams": M,
                        "n_batches": B, "xeng_id": xeng_id}}
    if desc:
        out = {"workload": desc, **out}
    return out


# The fixed bands of BASELINE.json configs[2..4], cut by frequency channel over the GPUs of one box
# (reference: coeff_generator.py:53 `ch = c + C * xeng_id`, prebeamform_reorder_test.py:72): (n_ants, n_chans of the
# whole band, n_samples, n_beams, the GPU counts the configuration is quoted for)
BANDS = {
    "c3": ("MeerKAT 4k mode: 64 antennas x 4096 channels x 256 samples, 64 beams", 64, 4096, 256, 64, (1, 2, 4, 8)),
    "c4": ("MeerKAT+ 32k mode: 80 antennas x 32768 channels x 256 samples, 32 beams", 80, 32768, 256, 32, (8,)),
    "c5": ("SKA-Mid scale: 197 antennas x 4096 channels x 256 samples, 256 beams", 197, 4096, 256, 256, (8,)),
}
_MODES = ("ms_per_step", "streaming_ms_per_step", "graph_ms_per_step", "graph_streaming_ms_per_step",
          "fp16_coeff_streaming_ms_per_step", "packed_ms_per_step", "packed_streaming_ms_per_step")


def strong_scaling(world, rank, dev, peak, steps, dist=None, emulate=(2, 4, 8)):
    """STRONG scaling: a fixed band cut over `world` GPUs (n_chans_per_gpu = n_chans / world, xeng_id = rank), the
    partitioning BASELINE.json configs[2..4] name.  Every rank times its own share (one heap and 8 heaps per launch;
    default launches, DCBF_FLAG_STREAMING, both replayed from a CUDA graph, streaming with the optional single fp16
    rounding of the coefficients, DCBF_FLAG_FP16_COEFF, and -- same beams bit for bit -- dcbf_fused_packed on coefficients
    packed once per delay model, default and streaming); the slowest rank's time is reported.
    `efficiency_vs_n1` = t(whole band on one GPU) / (world * t(share)), the whole band being timed in the same run.
    With one GPU the per-GPU shares of 2, 4 and 8 GPUs are timed on it instead (`emulated_shares`: the shards exchange
    nothing, so a share's kernel time does not depend on the other GPUs)."""
    import torch

    def timed(A, C, T, M, B, n_total, xid):
        r = measure_share(A, C, T, M, B, n_total, xid, steps, dev, peak)
        if world > 1:  # slowest rank, every mode
            t = torch.tensor([r[k] for k in _MODES], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            for k, v in zip(_MODES, t.tolist()):
                r[k] = v
        alg = r["algorithmic_bytes_per_launch"]
        for k in _MODES:  # (0: the mode does not exist for this shape -- packed coefficients at C5)
            r[k.replace("ms_per_step", "roofline_frac")] = alg / (r[k] / 1e3) / 1e9 / peak if r[k] else None
        r.pop("input_GBps", None)
        r.pop("beam_gsamples_per_s", None)
        return r

    out = {"note": strong_scaling.__doc__.split("\n\n")[0].replace("\n    ", " "), "n_gpus": world, "bands": {}}
    for name, (desc, A, C, T, M, quoted) in BANDS.items():
        if world > 1 and world not in quoted:
            continue
        band = {"workload": desc}
        for B in (1, 8):
            if B == 8 and name != "c3" and world == 1:
                continue  # the whole C4 / C5 band with 8 heaps does not need timing on one GPU
            key = f"b{B}"
            whole = timed(A, C, T, M, B, C, 0) if (name == "c3" or B == 1) else None
            band[key] = {"whole_band_on_one_gpu": whole}
            if world > 1:
                share = timed(A, C // world, T, M, B, C, rank)
                if whole is not None:
                    for k in _MODES:
                        share[k.replace("ms_per_step", "efficiency_vs_n1")] = whole[k] / (world * share[k]) if share[k] else None
                band[key]["share"] = share
            else:
                band[key]["emulated_shares"] = {}
                for n in emulate:
                    if n not in quoted:
                        continue
                    share = timed(A, C // n, T, M, B, C, n - 1)
                    if whole is not None:
                        for k in _MODES:
                            share[k.replace("ms_per_step", "efficiency_vs_n1")] = whole[k] / (n * share[k]) if share[k] else None
                    band[key]["emulated_shares"][f"n{n}"] = share
        out["bands"][name] = band
    return out


def measure_secondary(name, steps, dev, rank, world, peak):
    """Kernel-only figure of another BASELINE config on this GPU (see measure_share)."""
    desc, A, C, T, M, B = WORKLOADS[name]
    return measure_share(A, C, T, M, B, C * world, rank, steps, dev, peak, desc=f"{name}: {desc}")


def run_ours(args, wl) -> None:
    import torch
    import torch.distributed as dist

    from dpdk_dc_sand_b200 import _capi

    desc, A, C, T, M, B = wl
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved_stdout = None
    if world > 1:  # NCCL prints its version banner to fd 1: park stdout on stderr until the JSON line is due
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
    numa = None
    if world > 1 and not args.no_numa_bind:  # host staging buffers of the e2e leg next to this rank's GPU
        from dpdk_dc_sand_b200 import sharding

        numa = sharding.bind_to_device_numa(local)
        sys.stderr.write(f"[bench] rank {rank}: numa binding: {numa}\n")
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=dev)
    n_total = C * world
    flags = _capi.FLAG_FP16_COEFF if args.fp16_coeff else 0

    # synthetic inputs, generated on the device (seeded per rank); each rank owns channels [rank*C, (rank+1)*C)
    gen = torch.Generator(device=dev)
    gen.manual_seed(2021 + rank)
    samples = torch.randint(0, 256, (B, A, C, T, 2, 2), dtype=torch.uint8, device=dev, generator=gen)
    dv = torch.zeros((C, M, A, 4), dtype=torch.float32, device=dev)
    dv[..., 0] = (torch.rand((C, M, A), device=dev, generator=gen) * 32 - 16) * SAMPLE_PERIOD
    dv[..., 2] = (torch.rand((C, M, A), device=dev, generator=gen) * 2 - 1) * 3.14159265
    beams = torch.empty((B, 2, C, T // 16, 16, 2 * M), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()  # the inputs were generated on the default stream, the launches go to `stream`
    alg_bytes = _capi.fused_bytes(B, A, C, T, M)
    in_bytes = samples.numel()

    def step():
        _capi.fused(samples, dv, beams, B, A, C, n_total, T, M, rank, SAMPLE_PERIOD, flags, stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    _capi.fused_status()

    sampler = ClockSampler(local)
    sampler.start()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 3.0:  # NVML initialisation takes a moment
        time.sleep(0.01)
    # timed region: exactly K steps between two events on the launching stream (an event record between every
    # pair of launches would itself sit in the stream and lengthen each step by several microseconds)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _capi.launch_count()
    barrier()
    t_begin = time.time()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for i in range(args.steps):
            step()
        ev1.record(stream)
    stream.synchronize()
    barrier()
    t_end = time.time()
    launches = _capi.launch_count() - launches0
    clocks = sampler.window(t_begin, t_end)
    _capi.fused_status()
    total_ms = ev0.elapsed_time(ev1)
    # per-launch spread from a second, separately instrumented pass (not the figure reported as `value`)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(min(args.steps, 10) + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for i in range(len(ev) - 1):
            step()
            ev[i + 1].record(stream)
    stream.synchronize()
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(len(ev) - 1)]

    # ---- sustained: the same step after ~0.5 s of uninterrupted launches (a 1 kW part reaches its power cap) ----
    sustained = None
    if not args.no_sustained:
        time.sleep(0.2)
        t_s0 = time.time()
        while time.time() - t_s0 < 0.5:
            for _ in range(100):
                step()
            stream.synchronize()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_u0 = time.time()
        with torch.cuda.stream(stream):
            u0.record(stream)
            for _ in range(200):
                step()
            u1.record(stream)
        stream.synchronize()
        t_u1 = time.time()
        u_sec = u0.elapsed_time(u1) / 1e3 / 200
        sustained = {"ms_per_step": u_sec * 1e3, "algorithmic_GBps_per_gpu": alg_bytes / u_sec / 1e9,
                     "clocks": sampler.window(t_u0, t_u1),
                     "note": "200 steps timed after 0.5 s of continuous launches; reported beside the K-step figure"}
        time.sleep(0.3)
    sampler.stop()
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms_max = float(t.item())
    else:
        total_ms_max = total_ms
    sec_per_step = total_ms_max / 1e3 / args.steps

    # ---- streaming mode: consecutive heaps into alternating output buffers, DCBF_FLAG_STREAMING ----
    # (the next launch's CTAs may occupy SMs the previous launch has already left; every step still does all
    # of its work.  Reported beside the strictly serialised figure above, never instead of it.)
    streaming = None
    if not args.no_streaming:
        beams2 = torch.empty_like(beams)
        outs = (beams, beams2)
        sflags = flags | _capi.FLAG_STREAMING
        for i in range(3):
            _capi.fused(samples, dv, outs[i & 1], B, A, C, n_total, T, M, rank, SAMPLE_PERIOD, sflags, stream)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            s0.record(stream)
            for i in range(args.steps):
                _capi.fused(samples, dv, outs[i & 1], B, A, C, n_total, T, M, rank, SAMPLE_PERIOD, sflags, stream)
            s1.record(stream)
        stream.synchronize()
        barrier()
        _capi.fused_status()
        s_ms = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([s_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s_ms = float(t.item())
        s_sec = s_ms / 1e3 / args.steps
        streaming = {"ms_per_step": s_sec * 1e3, "value": world * in_bytes / s_sec / 1e9, "unit": UNIT,
                     "algorithmic_GBps_per_gpu": alg_bytes / s_sec / 1e9,
                     "same_result": bool(torch.equal(beams, beams2)),
                     "note": "DCBF_FLAG_STREAMING (programmatic dependent launch), 2 alternating output buffers"}
        del beams2

    # ---- single fp16 rounding of the coefficients (DCBF_FLAG_FP16_COEFF): one MMA pass instead of the hi + lo pair ----
    # (error <= 2^-12 per coefficient: inside north_star's 2^-10 sum|x| budget, outside the 1e-4 of the reference's own
    # unit tests, so it is an option and never the default.  At the board's power cap the default kernel follows the SM
    # clock, and what lowers that clock is the tensor cores' share of the power: this mode shows the figure without it.
    # Every rank runs it, rank 0's numbers are reported.)
    fp16c = None
    if not args.no_fp16_coeff and not args.fp16_coeff:
        beams_h = torch.empty_like(beams)
        hflags = flags | _capi.FLAG_FP16_COEFF

        def hstep():
            _capi.fused(samples, dv, beams_h, B, A, C, n_total, T, M, rank, SAMPLE_PERIOD, hflags, stream)

        for _ in range(3):
            hstep()
        stream.synchronize()
        time.sleep(0.3)
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            h0.record(stream)
            for _ in range(args.steps):
                hstep()
            h1.record(stream)
        stream.synchronize()
        _capi.fused_status()
        h_sec = h0.elapsed_time(h1) / 1e3 / args.steps
        # error against the default (hi + lo) result in units of the budget, on a slice of the channels
        cs = min(C, 64)
        xs = samples[:, :, :cs].to(torch.float32)
        bound = torch.sqrt(xs[..., 0] ** 2 + xs[..., 1] ** 2).sum(dim=1)  # [B][cs][T][pol]: sum over antennas of |x|
        bound = bound.permute(0, 3, 1, 2).reshape(B, 2, cs, T, 1)
        diff = (beams_h[:, :, :cs].reshape(B, 2, cs, T, 2 * M) - beams[:, :, :cs].reshape(B, 2, cs, T, 2 * M)).abs()
        err_in_budgets = float((diff / (bound * 2.0 ** -10)).max().item())
        del xs, bound, diff
        fp16c = {"ms_per_step": h_sec * 1e3, "value": world * in_bytes / h_sec / 1e9, "unit": UNIT,
                 "algorithmic_GBps_per_gpu": alg_bytes / h_sec / 1e9,
                 "max_abs_difference_from_default_in_budgets": err_in_budgets,
                 "budget": "2^-10 x sum over antennas of |x| per output sample (north_star); first %d channels" % cs,
                 "note": "DCBF_FLAG_FP16_COEFF: coefficients rounded once to fp16 (<= 2^-12 each), half the tensor-core work"}
        if not args.no_sustained:
            sampler2 = ClockSampler(local)
            sampler2.start()
            t_s0 = time.time()
            while time.time() - t_s0 < 0.5:
                for _ in range(100):
                    hstep()
                stream.synchronize()
            t_u0 = time.time()
            with torch.cuda.stream(stream):
                h0.record(stream)
                for _ in range(200):
                    hstep()
                h1.record(stream)
            stream.synchronize()
            t_u1 = time.time()
            hs_sec = h0.elapsed_time(h1) / 1e3 / 200
            fp16c["sustained"] = {"ms_per_step": hs_sec * 1e3, "algorithmic_GBps_per_gpu": alg_bytes / hs_sec / 1e9,
                                  "clocks": sampler2.window(t_u0, t_u1),
                                  "note": "200 steps timed after 0.5 s of continuous launches, like roofline.sustained"}
            sampler2.stop()
            time.sleep(0.3)
        del beams_h
        barrier()

    # ---- packed steering coefficients: the delay model evaluated once (dcbf_fused_pack_coeffs), every heap loads the
    # tile sets (dcbf_fused_packed).  Same beams bit for bit, same HBM bytes, a fraction of the SM-side work: the
    # full-precision figure of a deployment whose delay model changes at control-plane cadence.  Reported beside the
    # headline, which regenerates the coefficients per heap like the reference's sequence does. ----
    packed_blk = None
    n_packed = _capi.fused_packed_bytes(A, C, M, flags) if not args.no_packed else 0
    if n_packed:
        packed = torch.empty(n_packed, dtype=torch.uint8, device=dev)
        outs_p = (torch.empty_like(beams), torch.empty_like(beams))
        torch.cuda.synchronize()
        pk0, pk1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            _capi.fused_pack_coeffs(dv, packed, A, C, n_total, M, rank, SAMPLE_PERIOD, flags, stream)
            pk0.record(stream)
            for _ in range(5):
                _capi.fused_pack_coeffs(dv, packed, A, C, n_total, M, rank, SAMPLE_PERIOD, flags, stream)
            pk1.record(stream)
        stream.synchronize()

        def pstep(i=0, f=flags):
            _capi.fused_packed(samples, packed, outs_p[i & 1], B, A, C, n_total, T, M, rank, SAMPLE_PERIOD, f, stream)

        def timed_p(n, f, alternate):
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                q0.record(stream)
                for i in range(n):
                    pstep(i if alternate else 0, f)
                q1.record(stream)
            stream.synchronize()
            return q0.elapsed_time(q1) / 1e3 / n

        for _ in range(3):
            pstep()
        stream.synchronize()
        time.sleep(0.3)
        p_sec = timed_p(args.steps, flags, False)
        same_p = bool(torch.equal(outs_p[0], beams))
        time.sleep(0.3)
        timed_p(3, flags | _capi.FLAG_STREAMING, True)
        ps_sec = timed_p(args.steps, flags | _capi.FLAG_STREAMING, True)
        _capi.fused_status()
        packed_blk = {"ms_per_step": p_sec * 1e3, "value": world * in_bytes / p_sec / 1e9, "unit": UNIT,
                      "algorithmic_GBps_per_gpu": alg_bytes / p_sec / 1e9, "same_result_bit_for_bit": same_p,
                      "streaming_ms_per_step": ps_sec * 1e3, "streaming_algorithmic_GBps_per_gpu": alg_bytes / ps_sec / 1e9,
                      "pack_ms": pk0.elapsed_time(pk1) / 5, "packed_bytes": int(n_packed),
                      "api": "dcbf_fused_pack_coeffs once per delay model, dcbf_fused_packed per heap",
                      "note": "tile sets in the tensor cores' layout, as large as delay_vals: same HBM traffic, no phase / sin-cos "
                              "work per heap; every rank runs it, rank 0's numbers"}
        if not args.no_sustained:
            sampler3 = ClockSampler(local)
            sampler3.start()
            t_s0 = time.time()
            while time.time() - t_s0 < 0.5:
                for _ in range(100):
                    pstep()
                stream.synchronize()
            t_u0 = time.time()
            ps_sus = timed_p(200, flags, False)
            t_u1 = time.time()
            packed_blk["sustained"] = {"ms_per_step": ps_sus * 1e3, "algorithmic_GBps_per_gpu": alg_bytes / ps_sus / 1e9,
                                       "clocks": sampler3.window(t_u0, t_u1),
                                       "note": "200 steps timed after 0.5 s of continuous launches, like roofline.sustained"}
            sampler3.stop()
            time.sleep(0.3)
        del outs_p, packed
        barrier()

    # ---- requantised int8 output (dcbf_fused_q8): the SURVEY 8f-2 extension, reported beside the headline ----
    q8 = None
    if not args.no_q8:
        gains = torch.full((M,), 0.004, dtype=torch.float32, device=dev)
        out8 = torch.empty(beams.shape, dtype=torch.int8, device=dev)
        q8_bytes = _capi.fused_q8_bytes(B, A, C, T, M)
        for _ in range(3):
            _capi.fused_q8(samples, dv, gains, out8, B, A, C, n_total, T, M, rank, SAMPLE_PERIOD, flags, stream)
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            q0.record(stream)
            for _ in range(args.steps):
                _capi.fused_q8(samples, dv, gains, out8, B, A, C, n_total, T, M, rank, SAMPLE_PERIOD, flags, stream)
            q1.record(stream)
        stream.synchronize()
        barrier()
        _capi.fused_status()
        q_ms = q0.elapsed_time(q1)
        if world > 1:
            t = torch.tensor([q_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            q_ms = float(t.item())
        q_sec = q_ms / 1e3 / args.steps
        q8_packed_sec = None
        if _capi.fused_packed_bytes(A, C, M, flags) and not args.no_packed:  # the same on packed coefficients (bit-identical)
            pk = torch.empty(_capi.fused_packed_bytes(A, C, M, flags), dtype=torch.uint8, device=dev)
            out8p = torch.empty_like(out8)
            _capi.fused_pack_coeffs_q8(dv, gains, pk, A, C, n_total, M, rank, SAMPLE_PERIOD, flags, stream)
            for _ in range(3):
                _capi.fused_packed_q8(samples, pk, gains, out8p, B, A, C, n_total, T, M, rank, SAMPLE_PERIOD, flags, stream)
            stream.synchronize()
            time.sleep(0.3)
            with torch.cuda.stream(stream):
                q0.record(stream)
                for _ in range(args.steps):
                    _capi.fused_packed_q8(samples, pk, gains, out8p, B, A, C, n_total, T, M, rank, SAMPLE_PERIOD, flags, stream)
                q1.record(stream)
            stream.synchronize()
            _capi.fused_status()
            q8_packed_sec = q0.elapsed_time(q1) / 1e3 / args.steps
            q8_packed_same = bool(torch.equal(out8p, out8))
            del pk, out8p
        q8 = {"ms_per_step": q_sec * 1e3, "value": world * in_bytes / q_sec / 1e9, "unit": UNIT,
              "algorithmic_bytes_per_launch": q8_bytes, "algorithmic_GBps_per_gpu": q8_bytes / q_sec / 1e9,
              "note": "int8 beams = clip(rint(beam*gain)); output bytes / 4",
              "parity": "unpinned: the reference has no requantising path; checked against this repo's own oracle.requantise of the float64 beams"}
        if q8_packed_sec:
            q8["packed_coeffs"] = {"ms_per_step": q8_packed_sec * 1e3, "algorithmic_GBps_per_gpu": q8_bytes / q8_packed_sec / 1e9,
                                   "same_result_bit_for_bit": q8_packed_same,
                                   "api": "dcbf_fused_pack_coeffs_q8 once per delay model / gain set, dcbf_fused_packed_q8 per heap (rank 0's figure)"}
        del out8

    # ---- end to end through the host-buffer C-ABI call (pinned host arrays, H2D + kernel + D2H timed) ----
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty(samples.shape, dtype=torch.uint8, pin_memory=True)
        h_dv = torch.empty(dv.shape, dtype=torch.float32, pin_memory=True)
        h_out = torch.empty(beams.shape, dtype=torch.float32, pin_memory=True)
        h_in.copy_(samples)
        h_dv.copy_(dv)
        torch.cuda.synchronize()
        plan = _capi.HostPlan(B, A, C, n_total, T, M, rank, SAMPLE_PERIOD, flags, chunk_chans=max(1, C // 16), n_slots=3)
        n_in, n_dv, n_out = h_in.numpy(), h_dv.numpy(), h_out.numpy()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            plan.run(n_in, n_dv, n_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.run(n_in, n_dv, n_out)  # blocks until the beams are in host memory
        torch.cuda.synchronize()
        e2e_sec = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_sec], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_sec = float(t.item())
        # the host-path result must be the device-path result (same kernel, chunked): cheap identity check
        same = bool(torch.equal(h_out[:, :, : min(C, 8)], beams[:, :, : min(C, 8)].cpu()))
        # the same with the delay model resident on the device (it changes at control-plane cadence, not per heap):
        # per-step H2D = the voltages alone
        plan.set_delay_vals(n_dv)
        plan.run(n_in, None, n_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.run(n_in, None, n_out)
        torch.cuda.synchronize()
        r_sec = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([r_sec], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            r_sec = float(t.item())
        same_r = bool(torch.equal(h_out[:, :, : min(C, 8)], beams[:, :, : min(C, 8)].cpu()))
        plan.close()
        e2e = {"value": world * in_bytes / e2e_sec / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(in_bytes + dv.numel() * 4), "d2h_bytes_per_step": int(beams.numel() * 4),
               "ms_per_step": e2e_sec * 1e3, "steps": e2e_steps, "api": "dcbf_host_plan_run (pinned host arrays)",
               "matches_device_path": same,
               "resident_delay_model": {"value": world * in_bytes / r_sec / 1e9, "unit": UNIT, "ms_per_step": r_sec * 1e3,
                                        "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(beams.numel() * 4),
                                        "api": "dcbf_host_plan_set_delay_vals once, then dcbf_host_plan_run(samples, NULL, beams)",
                                        "matches_device_path": same_r}}
        if q8 is not None:
            h_out8 = torch.empty(beams.shape, dtype=torch.int8, pin_memory=True)
            n_out8 = h_out8.numpy()
            plan = _capi.HostPlan(B, A, C, n_total, T, M, rank, SAMPLE_PERIOD, flags, chunk_chans=max(1, C // 16), n_slots=3)
            plan.set_gains(gains.cpu().numpy())
            for _ in range(2):
                plan.run_q8(n_in, n_dv, n_out8)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                plan.run_q8(n_in, n_dv, n_out8)
            q_e2e = (time.perf_counter() - t0) / e2e_steps
            if world > 1:
                t = torch.tensor([q_e2e], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                q_e2e = float(t.item())
            plan.set_delay_vals(n_dv)
            plan.run_q8(n_in, None, n_out8)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                plan.run_q8(n_in, None, n_out8)
            q_res = (time.perf_counter() - t0) / e2e_steps
            if world > 1:
                t = torch.tensor([q_res], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                q_res = float(t.item())
            plan.close()
            q8["e2e"] = {"value": world * in_bytes / q_e2e / 1e9, "unit": UNIT, "ms_per_step": q_e2e * 1e3,
                         "h2d_bytes_per_step": int(in_bytes + dv.numel() * 4), "d2h_bytes_per_step": int(beams.numel()),
                         "resident_delay_model": {"value": world * in_bytes / q_res / 1e9, "unit": UNIT,
                                                  "ms_per_step": q_res * 1e3, "h2d_bytes_per_step": int(in_bytes),
                                                  "d2h_bytes_per_step": int(beams.numel())}}
            del h_out8
        del h_in, h_dv, h_out

    gather = None
    if world > 1 and not args.no_gather:
        # the path's only (optional) collective, outside the compute timing: every rank's beams to rank 0 over NCCL
        from dpdk_dc_sand_b200 import sharding

        shard = sharding.plan(n_total, world=world, rank=rank)
        _capi.fused(samples, dv, beams, B, A, C, n_total, T, M, rank, SAMPLE_PERIOD, flags)
        torch.cuda.synchronize()
        local_sum = beams.double().sum()
        total = local_sum.clone()
        dist.all_reduce(total)
        times = []
        full = None
        for _ in range(2):
            del full
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            full = sharding.gather_beams(beams, shard, dst=0)
            e1.record()
            torch.cuda.synchronize()
            t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
            times.append(float(t_ms.item()))
        ok = True
        if rank == 0:
            ok = tuple(full.shape)[2] == n_total and abs(float(full.double().sum().item()) - float(total.item())) <= 1e-6 * abs(float(total.item())) + 1e-3 \
                and torch.equal(full[:, :, :C], beams)
        del full
        torch.cuda.empty_cache()
        gbytes = (world - 1) * beams.numel() * 4
        gather = {"ms": min(times), "bytes_received_by_rank0": gbytes, "GBps": gbytes / (min(times) / 1e3) / 1e9,
                  "checksum_ok": bool(ok), "api": "sharding.gather_beams (torch.distributed gather, nccl)",
                  "note": "optional; not part of value / e2e"}

    peak, peak_src = _peaks()
    strong = None
    if not args.no_strong and args.workload == "c3":
        del samples, dv, beams
        torch.cuda.empty_cache()
        strong = strong_scaling(world, rank, dev, peak, max(args.steps, 30), dist if world > 1 else None)
        samples = dv = beams = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    secondary = None
    if world == 1 and not args.no_secondary and args.workload == "c3":
        torch.cuda.empty_cache()
        secondary = {"c2": measure_secondary("c2", max(args.steps, 40), dev, rank, world, peak),
                     "c2_b8": measure_secondary("c2_b8", max(args.steps, 20), dev, rank, world, peak)}
    mean_launch_s = total_ms / 1e3 / args.steps  # this rank's average launch duration over the timed region
    achieved = alg_bytes / mean_launch_s / 1e9
    if q8 is not None:
        q8["roofline_frac"] = q8["algorithmic_GBps_per_gpu"] / peak
        if "packed_coeffs" in q8:
            q8["packed_coeffs"]["roofline_frac"] = q8["packed_coeffs"]["algorithmic_GBps_per_gpu"] / peak
    if packed_blk is not None:
        packed_blk["roofline_frac"] = packed_blk["algorithmic_GBps_per_gpu"] / peak
        packed_blk["streaming_roofline_frac"] = packed_blk["streaming_algorithmic_GBps_per_gpu"] / peak
        if "sustained" in packed_blk:
            packed_blk["sustained"]["roofline_frac"] = packed_blk["sustained"]["algorithmic_GBps_per_gpu"] / peak
    if fp16c is not None:
        fp16c["roofline_frac"] = fp16c["algorithmic_GBps_per_gpu"] / peak
        if "sustained" in fp16c:
            fp16c["sustained"]["roofline_frac"] = fp16c["sustained"]["algorithmic_GBps_per_gpu"] / peak
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": _traffic(args.workload),
                "traffic_source": "committed ncu capture (profiles/fused_traffic.json), not measured in this run",
                "kernel": "fused_beamform_kernel",
                "frac_sustained": (alg_bytes / (sustained["ms_per_step"] / 1e3) / 1e9 / peak) if sustained else None,
                "sustained": sustained,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_us_mean": mean_launch_s * 1e6,
                "launch_us_min": min(per_launch_ms) * 1e3, "launch_us_with_event_between_launches": statistics.mean(per_launch_ms) * 1e3,
                "kernel_us_under_ncu": _ncu_kernel_us(args.workload), "kernel_us_under_ncu_source": "committed ncu capture",
                "peak_source": peak_src,
                "tensor_tflops_real_expanded": B * 2 * C * T * 8 * A * M / mean_launch_s / 1e12}
    cpu = None
    if world == 1 and not args.no_cpu:
        sample = min(C, 256)
        rate, threads, times = cpu_port_rate(A, M, T, n_total, sample, repeats=5, budget_s=20.0)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{sample} of {C} channels x {len(times)} repeats (median), oracle numpy port "
                         f"(transpose + float64 phase + float32 BLAS matmul), {getattr(cpu_port_rate, 'how', '')}; host has "
                         f"{os.cpu_count()} logical cores"}
    line = {
        "metric": METRIC, "value": world * in_bytes / sec_per_step / 1e9, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16",  # tcgen05 kind::f16 operands (u8 voltages exact, coefficients fp16 hi+lo pairs), f32 accumulate
        "data": "synthetic",
        "config": bench_config(args.workload, wl, world),
        "arithmetic": "u8 voltages exact in f16; coefficients as f16 " + ("single rounding" if args.fp16_coeff else "hi+lo pair (~2^-24)") + "; f32 accumulate in TMEM; f32 beams",
        "tiling": dict(zip(("kb_count", "nt", "nt_count"), _capi.fused_tiling(A, M, flags))),
        "beam_gsamples_per_s": world * B * 2 * C * T * M / sec_per_step / 1e9,
        "algorithmic_GBps": world * alg_bytes / sec_per_step / 1e9,
        "roofline": roofline, "cpu_baseline": cpu, "cpu_baseline_verbatim": _verbatim_cpu(), "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "sustained": sustained, "streaming": streaming, "fp16_coeff": fp16c, "packed_coeffs": packed_blk, "q8_output": q8, "other_workloads": secondary, "gather": gather,
        "strong_scaling": strong,
        "pcie": {"note": "e2e is bounded by the host link: pinned copies measured on this pool (tools/bench_standalone.py) "
                         "reach 55.5 GB/s H2D, 57.3 GB/s D2H alone and 49.9 GB/s each way when both directions run "
                         "at once; the e2e step moves h2d_bytes_per_step up and d2h_bytes_per_step down",
                 "e2e_d2h_GBps": (e2e["d2h_bytes_per_step"] / (e2e["ms_per_step"] / 1e3) / 1e9) if e2e else None},
    }
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["c1", "c2", "c3"], default="c3")
    ap.add_argument("--fp16-coeff", action="store_true", help="single fp16 coefficient rounding (DCBF_FLAG_FP16_COEFF)")
    ap.add_argument("--no-packed", action="store_true", help="skip the packed_coeffs block (dcbf_fused_pack_coeffs / dcbf_fused_packed)")
    ap.add_argument("--no-fp16-coeff", action="store_true", help="skip the fp16_coeff block (the optional single-rounding mode beside the default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-streaming", action="store_true")
    ap.add_argument("--no-sustained", action="store_true", help="skip the power-capped sustained-load measurement")
    ap.add_argument("--no-q8", action="store_true", help="skip the int8-output extension measurement")
    ap.add_argument("--no-secondary", action="store_true", help="skip the extra c2 (BASELINE configs[1]) measurement")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block (fixed bands cut over the GPUs)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: leave the rank's CPU affinity alone")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip timing the optional beam-output gather to rank 0")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
