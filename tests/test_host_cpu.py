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
