"""CPU-only checks of the drop-in boundary: the C-ABI library exports what include/dcbf.h declares, argument
validation works without a device, the operator templates reproduce the reference's shape algebra, and the
channel sharding (rank == xeng_id) composes to the full-band result over a 2-rank gloo group.
No CUDA compute is called here; the oracle is only the checker."""
import ctypes
import math
import os
import re
import socket

import numpy as np
import pytest

from oracle import beamform_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dcbf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcbf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from dpdk_dc_sand_b200 import _capi

    lib = _capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in dcbf.h but not exported"
        assert name in _capi.SIGNATURES, f"{name} has no ctypes prototype"
    assert sorted(_capi.SIGNATURES) == declared
    assert lib.dcbf_version() == 1


def test_argument_validation_needs_no_device():
    from dpdk_dc_sand_b200 import _capi

    lib = _capi.load()
    assert lib.dcbf_reorder(None, None, 1, 1, 1, 16, None) == _capi.ERR_INVALID_ARG
    assert lib.dcbf_fused(None, None, None, 1, 1, 1, 1, 16, 1, 0, 1e-9, 0, None) == _capi.ERR_INVALID_ARG
    buf = (ctypes.c_uint8 * 64)()
    p = ctypes.addressof(buf)
    p += (-p) % 16
    assert lib.dcbf_reorder(p, p, 1, 1, 1, 24, None) == _capi.ERR_INVALID_ARG  # T % 16 != 0
    assert lib.dcbf_reorder(p + 1, p, 1, 1, 1, 16, None) == _capi.ERR_INVALID_ARG  # misaligned
    assert lib.dcbf_coeffs(p, p, 1, 2, 1, 1, 1, 1, 0, -1.0, None) == _capi.ERR_INVALID_ARG  # sample_period <= 0
    for code in (_capi.OK, _capi.ERR_INVALID_ARG, _capi.ERR_UNSUPPORTED, _capi.ERR_CUDA, _capi.ERR_NO_DEVICE,
                 _capi.ERR_TIMEOUT):
        assert lib.dcbf_strerror(code)
    with pytest.raises(ValueError):
        _capi.check(_capi.ERR_INVALID_ARG, "x")
    with pytest.raises(_capi.DcbfError):
        _capi.check(_capi.ERR_TIMEOUT, "x")


def test_algorithmic_bytes_and_tiling():
    from dpdk_dc_sand_b200 import _capi

    # SURVEY.md section 8(d) / BASELINE.md section 3
    assert _capi.fused_bytes(1, 4, 64, 256, 4) == 1_327_104
    assert _capi.fused_bytes(1, 64, 1024, 256, 16) == 150_994_944
    assert _capi.fused_bytes(1, 64, 4096, 256, 64) == 1_610_612_736
    assert _capi.fused_bytes(1, 80, 32768, 256, 32) == 8_321_499_136
    assert _capi.fused_bytes(1, 197, 4096, 256, 256) == 8_426_356_736
    for a, m in [(4, 4), (64, 16), (64, 64), (80, 32), (197, 256), (5, 3), (300, 2)]:
        for flags in (0, _capi.FLAG_FP16_COEFF):
            kb, nt, ntc = _capi.fused_tiling(a, m, flags)
            parts = 1 if flags else 2
            assert kb == -(-a // 32) and nt % 16 == 0 and 16 <= nt <= 128
            assert nt * ntc >= 2 * m and kb * parts * nt * 128 <= 64 * 1024
            # packed steering coefficients: one tile set per channel where a whole one exists, none for K-streamed shapes
            want = 7 * kb * parts * nt * 128 if ntc == 1 else 0
            assert _capi.fused_packed_bytes(a, 7, m, flags) == want
    # at C3 a channel's tile set is exactly as large as its delay_vals (64 KiB): the hot path's HBM traffic is unchanged
    assert _capi.fused_packed_bytes(64, 4096, 64) == 4096 * 64 * 64 * 16 == 4096 * 65536
    assert _capi.fused_packed_bytes(64, 4096, 64, _capi.FLAG_FP16_COEFF) == 4096 * 32768
    assert _capi.fused_packed_bytes(197, 512, 256) == 0 and _capi.fused_packed_bytes(0, 1, 1) == 0


def test_templates_reproduce_reference_shape_algebra():
    import dpdk_dc_sand_b200

    dpdk_dc_sand_b200.install_dropin()
    from beamforming.beamform_op_sequence import OpSequenceTemplate
    from beamforming.coeff_generator import CoeffGeneratorTemplate
    from beamforming.matrix_multiply import MatrixMultiplyTemplate
    from beamforming.prebeamform_reorder import PreBeamformReorderTemplate

    b, a, c, t, m, n = 3, 79, 103, 256, 2, 32768
    r = PreBeamformReorderTemplate(None, a, c, t, b)
    assert tuple(d.size for d in r.inputDataShape) == (b, a, c, t, 2, 2)
    assert tuple(d.size for d in r.outputDataShape) == (b, 2, c, 16, 16, a, 2)
    assert all(d.exact for d in r.inputDataShape + r.outputDataShape)
    assert r.matrix_size == a * c * t * 2 and r.n_blocks_x == math.ceil(r.matrix_size / 1024)
    with pytest.raises(ValueError):
        PreBeamformReorderTemplate(None, a, c, 24, b)
    cg = CoeffGeneratorTemplate(None, b, 2, c, n, 16, 16, a, m, 0, orc.SAMPLE_PERIOD)
    assert tuple(d.size for d in cg.delay_vals_data_dimensions) == (c, m, a, 4)
    assert tuple(d.size for d in cg.coeff_data_dimensions) == (b, 2, c, 2 * a, 2 * m)
    mm = MatrixMultiplyTemplate(None, a, c, t, m, b)
    assert tuple(d.size for d in mm.input_data_dimensions) == (b, 2, c, 16, 16, a, 2)
    assert tuple(d.size for d in mm.output_data_dimensions) == (b, 2, c, 16, 16, 2 * m)
    assert tuple(d.size for d in mm.coeff_data_dimensions) == (b, 2, c, 2 * a, 2 * m)

    class _Queue:  # no device: slot wiring only
        context = None

    op = OpSequenceTemplate(None, b, 2, c, n, 16, 16, a, m, 0, orc.SAMPLE_PERIOD, t).instantiate(_Queue())
    assert set(op.slots) == {"bufin_delay_vals", "bufint_coeff", "bufin_reorder", "bufint_data", "bufout_mult"}
    assert op.slots["bufin_reorder"].shape == (b, a, c, t, 2, 2)
    assert op.slots["bufout_mult"].shape == (b, 2, c, 16, 16, 2 * m)
    assert op.slots["bufout_mult"].dtype == np.float32 and op.slots["bufin_reorder"].dtype == np.uint8
    assert op.prebeamform_reorder.slots["outReordered"] is not None and op.beamform_mult.slots["inCoeffs"] is not None
    with pytest.raises(ValueError):
        op()  # nothing bound, and no device: must refuse, never fall back to a CPU path


def test_sharding_plan():
    from dpdk_dc_sand_b200 import sharding

    s = sharding.plan(4096, world=8, rank=3)
    assert (s.xeng_id, s.n_channels_per_stream, s.first_channel) == (3, 512, 1536)
    assert s.channels == slice(1536, 2048)
    with pytest.raises(ValueError):
        sharding.plan(4097, world=8, rank=0)
    with pytest.raises(ValueError):
        sharding.plan(4096, world=8, rank=8)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from dpdk_dc_sand_b200 import sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, a, n, t, m = 1, 6, 8, 32, 3
        x = orc.make_samples(b, a, n, t, seed=5)            # every rank builds the same full-band input
        dv = orc.make_delay_vals_random(n, m, a, seed=6)
        shard = sharding.plan(n)                             # from the RANK / WORLD_SIZE environment
        xl, dvl = sharding.local_samples(x, shard), sharding.local_delay_vals(dv, shard)
        assert xl.shape == (b, a, n // world, t, 2, 2) and xl.flags["C_CONTIGUOUS"]
        # the rank-local computation an X-engine performs (oracle stands in for the GPU kernel on CPU)
        local = orc.beamform_pipeline(xl, dvl, shard.n_channels, shard.xeng_id, orc.SAMPLE_PERIOD)
        full = sharding.gather_beams(torch.from_numpy(local), shard, dst=0)
        if rank == 0:
            ref = orc.beamform_pipeline(x, dv, n, 0, orc.SAMPLE_PERIOD)  # one engine owning the whole band
            q.put(("ok", float(np.abs(full.numpy() - ref).max()), tuple(full.shape)))
        else:
            assert full is None
    except Exception as exc:  # pragma: no cover
        if rank == 0:
            q.put(("error", repr(exc), None))
        raise
    finally:
        dist.destroy_process_group()


def test_channel_sharding_composes_to_full_band_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    status, err, shape = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert status == "ok", err
    assert shape == (1, 2, 8, 2, 16, 6)
    assert err < 1e-9  # same float64 arithmetic, channel offset carried by xeng_id


def test_ingest_ring_assembles_heaps_in_any_order():
    """The ingest stage (host C++, no GPU): heaps placed by (timestamp, feng_id) in shuffled order come out as
    [B][A][C][T][2][2] chunks in time order; a lost heap is zero-filled and flagged once newer data forces its
    chunk out; late and duplicate heaps are counted, never written."""
    from dpdk_dc_sand_b200 import _capi

    n_chunks, B, A, C_, T, step = 4, 2, 3, 2, 16, 8192
    ing = _capi.Ingest(n_chunks, B, A, C_, T, step, pinned=False)
    rng = np.random.default_rng(11)
    n_heaps_t = 8  # 4 chunks of 2 heaps
    truth = rng.integers(0, 256, (n_heaps_t, A, C_, T, 2, 2), dtype=np.uint8)
    lost = (3, 1)  # (heap number, antenna) that never arrives
    popped = []

    def drain(flush=False):
        while True:
            got = ing.pop(flush)
            if got is None:
                return
            samples, ts, present = got
            popped.append((ts, samples.copy(), present.copy()))
            ing.release(samples)

    for pair in range(0, n_heaps_t, 4):  # feed two chunks' worth at a time, shuffled inside
        order = [(h, a) for h in range(pair, min(pair + 4, n_heaps_t)) for a in range(A) if (h, a) != lost]
        rng.shuffle(order)
        for h, a in order:
            assert ing.heap(h * step, a, truth[h, a])
        drain()
    assert not ing.heap(0, 0, truth[0, 0])            # chunk 0 is long gone: dropped, not written
    assert ing.stats()["late_or_dropped"] == 1
    with pytest.raises(ValueError):
        ing.heap(step // 2, 0, truth[0, 0])           # not a heap boundary
    drain(flush=True)
    assert [p[0] for p in popped] == [0, 2 * step, 4 * step, 6 * step]  # time order
    for k, (ts, samples, present) in enumerate(popped):
        want = truth[2 * k:2 * k + 2].copy()
        exp_present = np.ones((B, A), bool)
        if k == 1:
            want[1, 1] = 0
            exp_present[1, 1] = False
        np.testing.assert_array_equal(present, exp_present)
        np.testing.assert_array_equal(samples, want)
    ing.close()


def _spead_packet(heap_cnt, heap_size, heap_offset, payload, items):
    """One SPEAD-64-48 packet: header, the four standard immediates, then ``items`` = [(id, immediate?, value)].
    Layout per the SPEAD specification as spead2's Flavour(4, 64, 48) emits it (fgpu_send_prototype.py:18)."""
    import struct

    ptrs = [(0x0001, True, heap_cnt), (0x0002, True, heap_size), (0x0003, True, heap_offset),
            (0x0004, True, len(payload))] + list(items)
    out = bytes([0x53, 0x04, 2, 6, 0, 0]) + struct.pack(">H", len(ptrs))
    for ident, immediate, value in ptrs:
        out += struct.pack(">Q", (int(immediate) << 63) | (ident << 48) | (value & 0xFFFFFFFFFFFF))
    return out + bytes(payload)


def test_ingest_accepts_raw_spead_packets():
    """dcbf_ingest_packet: heaps arrive as SPEAD-64-48 packets (timestamp 0x1600, feng_id 0x4101, frequency 0x4103,
    feng_raw 0x4300), several packets per heap, interleaved between antennas; only the first packet of a heap
    carries the item pointers (spead2's default).  The chunk comes out identical to heap-wise ingest."""
    from dpdk_dc_sand_b200 import _capi

    B, A, C_, T, step = 2, 3, 4, 16, 4096
    heap_bytes = C_ * T * 4
    ing = _capi.Ingest(4, B, A, C_, T, step, pinned=False)
    ing.set_frequency(8)
    rng = np.random.default_rng(5)
    truth = rng.integers(0, 256, (B, A, C_, T, 2, 2), dtype=np.uint8)
    pkt_payload = 96  # heap_bytes = 256 -> packets of 96, 96, 64 bytes
    streams = []
    for b in range(B):
        for a in range(A):
            raw = truth[b, a].tobytes()
            cnt = (b * A + a) * 7 + 3
            pkts = []
            for off in range(0, heap_bytes, pkt_payload):
                first = off == 0
                items = [(0x1600, True, b * step), (0x4101, True, a), (0x4103, True, 8), (0x4300, False, 0)] if first else []
                pkts.append(_spead_packet(cnt, heap_bytes, off, raw[off:off + pkt_payload], items))
            streams.append(pkts)
    # a heap of another sub-band, a descriptor heap and garbage: refused without touching the chunk
    other = _spead_packet(999, heap_bytes, 0, bytes(heap_bytes), [(0x1600, True, 0), (0x4101, True, 0), (0x4103, True, 12),
                                                                  (0x4300, False, 0)])
    assert not ing.packet(other)
    assert not ing.packet(_spead_packet(1000, 0, 0, b"", [(0x0005, False, 0)]))
    with pytest.raises(ValueError):
        ing.packet(b"\x53\x04\x02\x06\x00\x00\x00\x09short")
    assert ing.stats()["bad"] == 1
    # round-robin over the heaps (packets of different antennas interleave), later packets of two heaps swapped
    rounds = [[s[i] for s in streams if i < len(s)] for i in range(3)]
    rounds[1], rounds[2] = rounds[2], rounds[1]
    assert ing.pop() is None
    for r in rounds:
        for pkt in r:
            assert ing.packet(pkt)
    got = ing.pop()
    assert got is not None
    samples, ts, present = got
    assert ts == 0 and present.all()
    np.testing.assert_array_equal(samples, truth)
    ing.release(samples)
    # a packet that overtakes its heap's first packet cannot be placed; a sender without feng_id uses the default
    late = _spead_packet(5000, heap_bytes, 96, bytes(96), [])
    assert not ing.packet(late)
    solo = _spead_packet(5001, heap_bytes, 0, bytes(range(256)), [(0x1600, True, 2 * step), (0x4300, False, 0)])
    assert ing.packet(solo, default_feng_id=1)
    samples, ts, present = ing.pop(flush=True)
    assert ts == 2 * step and present.sum() == 1 and present[0, 1]
    np.testing.assert_array_equal(samples[0, 1].reshape(-1), np.arange(256, dtype=np.uint8))
    ing.close()


def test_ingest_duplicate_and_retransmitted_packets_do_not_complete_a_heap_early():
    """A repeated packet must not count twice towards its heap: the heap is complete only when every byte range has
    arrived (a byte count alone would report it present while the missing packet's bytes are still those of the slot's
    previous chunk, and would then drop the real packet as a duplicate).  A 48-bit feng_id beyond the antenna count is
    refused instead of being narrowed to a valid one."""
    from dpdk_dc_sand_b200 import _capi

    B, A, C_, T, step = 1, 2, 4, 16, 4096
    heap_bytes = C_ * T * 4  # 256: packets of 96, 96, 64 bytes
    ing = _capi.Ingest(4, B, A, C_, T, step, pinned=False)
    rng = np.random.default_rng(9)
    truth = rng.integers(1, 256, (B, A, C_, T, 2, 2), dtype=np.uint8)

    def packets(a, cnt, data):
        raw = data.tobytes()
        out = []
        for off in range(0, heap_bytes, 96):
            items = [(0x1600, True, 0), (0x4101, True, a), (0x4300, False, 0)] if off == 0 else []
            out.append(_spead_packet(cnt, heap_bytes, off, raw[off:off + 96], items))
        return out

    p0, p1 = packets(0, 11, truth[0, 0]), packets(1, 12, truth[0, 1])
    for pkt in p1:
        assert ing.packet(pkt)
    assert ing.packet(p0[0])
    assert ing.packet(p0[1])
    assert ing.packet(p0[1])      # duplicate: 96 + 96 + 96 bytes "received", but [192, 256) is still missing
    assert ing.packet(p0[0])      # retransmission of the first packet
    assert ing.stats()["duplicate"] == 2
    assert ing.pop() is None      # not complete yet
    assert ing.packet(p0[2])      # the real last packet is NOT dropped as a duplicate
    samples, ts, present = ing.pop()
    assert present.all()
    np.testing.assert_array_equal(samples, truth)
    ing.release(samples)
    wide = _spead_packet(13, heap_bytes, 0, bytes(96), [(0x1600, True, step), (0x4101, True, 0x100000001), (0x4300, False, 0)])
    with pytest.raises(ValueError):
        ing.packet(wide)
    ing.close()


def test_numa_binding_helper_is_optional():
    """sharding.bind_to_device_numa is an optimisation for one-process-per-GPU hosts: without NVML / a GPU it reports
    None and leaves the process alone."""
    import os

    from dpdk_dc_sand_b200 import sharding

    before = os.sched_getaffinity(0)
    out = sharding.bind_to_device_numa(0)
    assert out is None or isinstance(out, str)
    if out is None:
        assert os.sched_getaffinity(0) == before


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores; needs no GPU) prints exactly one JSON line with
    the contract keys; the product arm refuses to run without a CUDA device instead of falling back."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    proc = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                           "--workload", "c2"], capture_output=True, text=True, timeout=600, env=env)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    ours = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, env=env)
    assert ours.returncode != 0 and "CUDA" in (ours.stderr + ours.stdout)


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/dcbf.h is the drop-in boundary: it must compile as C99 (no C++ in any signature) and a C program
    must link against libdcbf.so and call the entry points that need no GPU."""
    import shutil
    import subprocess

    from dpdk_dc_sand_b200 import _capi

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    _capi.load()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "dpdk_dc_sand_b200", "lib")
    src = tmp_path / "smoke.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "dcbf.h"
int main(void) {
    int kb = 0, nt = 0, ntc = 0;
    if (dcbf_version() != DCBF_VERSION) return 1;
    if (strlen(dcbf_strerror(DCBF_ERR_TIMEOUT)) == 0) return 2;
    if (dcbf_fused_bytes(1, 64, 4096, 256, 64) != 1610612736LL) return 3;
    dcbf_fused_tiling(64, 64, 0, &kb, &nt, &ntc);
    if (kb != 2 || nt != 128 || ntc != 1) return 4;
    if (dcbf_reorder(NULL, NULL, 1, 1, 1, 16, NULL) != DCBF_ERR_INVALID_ARG) return 5;
    dcbf_ingest_t ing = NULL;
    if (dcbf_ingest_create(&ing, 2, 1, 2, 2, 16, 1024, 0) != DCBF_OK) return 6;
    if (dcbf_ingest_destroy(ing) != DCBF_OK) return 7;
    printf("ok\n");
    return 0;
}
''')
    exe = tmp_path / "smoke"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src), "-L", lib_dir,
                    "-ldcbf", "-Wl,-rpath," + lib_dir, "-o", str(exe)], check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout, out.stderr)
