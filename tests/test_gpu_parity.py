"""Parity of the CUDA path (through the operator API -> ctypes -> libdcbf C ABI) against the oracle.

Mirrors the reference's four GPU tests (beamformer/unit_test/*_test.py): same calling protocol, same
synthetic inputs (rng seed 2021; uniform delay 5*Ts / phase pi/2), plus the cases those tests miss
(non-uniform delays, M > 2, xeng_id > 0, signed input, ragged antenna counts, partial time tiles).
Tolerances: reorder bit-exact; coefficients <= 1e-6 of a float64 evaluation (and exact float32 equality on the
reference's own uniform inputs); beams max|err| <= 2^-10 * sum_a |x_a| per output sample, and the
reference's rtol=atol=1e-4 on its own test inputs.
"""
import math
import os

import numpy as np
import pytest

from oracle import beamform_oracle as orc

pytestmark = pytest.mark.gpu
TS = orc.SAMPLE_PERIOD


@pytest.fixture(scope="module")
def dropin():
    import dpdk_dc_sand_b200

    dpdk_dc_sand_b200.install_dropin()
    from katsdpsigproc import accel

    ctx = accel.create_some_context(device_filter=lambda d: d.is_cuda, interactive=False)
    return ctx, ctx.create_command_queue()


def _budget(x, signed=False):
    return 2.0 ** -10 * orc.beamform_abs_bound(orc.reorder(x), signed_input=signed)[..., None]


# reference grid: n_ants (test_parameters.py:19) with C = N // A // 4 (prebeamform_reorder_test.py:72), B = 3
REF_ANTS = [4, 8, 16, 32, 64, 79, 80, 84, 130, 192, 256, 5, 23, 61, 19]


@pytest.mark.parametrize("n_ants", REF_ANTS)
def test_prebeamform_reorder(dropin, n_ants):
    from beamforming.prebeamform_reorder import PreBeamformReorderTemplate

    ctx, queue = dropin
    n_batches, n_channels, n_samples = 3, 1024, 256
    c = max(1, n_channels // n_ants // 4)
    op = PreBeamformReorderTemplate(ctx, n_ants, c, n_samples, n_batches).instantiate(queue)
    op.ensure_all_bound()
    buf_in, buf_out = op.buffer("inSamples"), op.buffer("outReordered")
    host_in = buf_in.empty_like()
    rng = np.random.default_rng(seed=2021)
    host_in[:] = rng.uniform(0, 255, host_in.shape).astype(np.uint8)  # prebeamform_reorder_test.py:100-106
    buf_in.set(queue, host_in)
    op()
    got = buf_out.get(queue)
    np.testing.assert_array_equal(got, orc.reorder(np.asarray(host_in)))


@pytest.mark.parametrize("shape", [(1, 3, 2, 16), (2, 7, 5, 48), (1, 64, 3, 512), (1, 300, 1, 32)])
def test_prebeamform_reorder_ragged(dropin, shape):
    from dpdk_dc_sand_b200 import _capi
    import torch

    b, a, c, t = shape
    x = orc.make_samples(b, a, c, t, seed=5)
    dx = torch.from_numpy(x).cuda()
    out = torch.zeros((b, 2, c, t // 16, 16, a, 2), dtype=torch.uint8, device="cuda")
    _capi.reorder(dx, out, b, a, c, t)
    np.testing.assert_array_equal(out.cpu().numpy(), orc.reorder(x))


@pytest.mark.parametrize("n_ants,n_channels", [(4, 1024), (64, 4096), (79, 32768), (23, 1024)])
def test_coeff_generator_reference_inputs_exact(dropin, n_ants, n_channels):
    """beamform_coeff_test.py: uniform delay 5*Ts, phase pi/2 -> exact float32 equality with the CPU values."""
    from beamforming.coeff_generator import CoeffGeneratorTemplate

    ctx, queue = dropin
    b, p, m, xid = 3, 2, 2, 0
    c = max(1, n_channels // n_ants // 4)
    op = CoeffGeneratorTemplate(ctx, b, p, c, n_channels, 16, 16, n_ants, m, xid, TS).instantiate(queue)
    op.ensure_all_bound()
    dv = orc.make_delay_vals_uniform(c, m, n_ants)
    op.buffer("delay_vals").set(queue, dv)
    op()
    got = op.buffer("outCoeffs").get(queue)
    np.testing.assert_array_equal(got, orc.steering_coeffs(dv, b, p, c, n_channels, n_ants, m, xid, TS))


@pytest.mark.parametrize("n_ants,n_beams,xid", [(5, 3, 1), (64, 16, 0), (80, 32, 7), (197, 9, 2)])
def test_coeff_generator_random(dropin, n_ants, n_beams, xid):
    from beamforming.coeff_generator import CoeffGeneratorTemplate

    ctx, queue = dropin
    b, p, c, n = 1, 2, 6, 4096
    op = CoeffGeneratorTemplate(ctx, b, p, c, n, 16, 16, n_ants, n_beams, xid, TS).instantiate(queue)
    op.ensure_all_bound()
    dv = orc.make_delay_vals_random(c, n_beams, n_ants, seed=11)
    op.buffer("delay_vals").set(queue, dv)
    op()
    got = op.buffer("outCoeffs").get(queue).astype(np.float64)
    ref = orc.steering_coeffs(dv, b, p, c, n, n_ants, n_beams, xid, TS, out_dtype=np.float64)
    assert np.abs(got - ref).max() <= 1e-6  # unit-magnitude phasors: absolute == relative to |coeff|
    # block structure
    np.testing.assert_array_equal(got[..., 1::2, 0::2], -got[..., 0::2, 1::2])
    np.testing.assert_array_equal(got[..., 1::2, 1::2], got[..., 0::2, 0::2])


@pytest.mark.parametrize("n_ants,n_beams,signed", [(4, 2, False), (64, 16, False), (23, 5, True), (130, 2, False),
                                                   (80, 32, True), (64, 64, False)])
def test_matrix_multiply(dropin, n_ants, n_beams, signed):
    from beamforming.matrix_multiply import MatrixMultiplyTemplate

    ctx, queue = dropin
    b, c, t = 2, 3, 64
    op = MatrixMultiplyTemplate(ctx, n_ants, c, t, n_beams, b).instantiate(queue)
    op.signed_input = signed
    op.ensure_all_bound()
    x = orc.make_samples(b, n_ants, c, t, seed=3)
    re = orc.reorder(x)
    co = np.random.default_rng(4).standard_normal((b, 2, c, 2 * n_ants, 2 * n_beams)).astype(np.float32)
    op.buffer("inData").set(queue, re)
    op.buffer("inCoeffs").set(queue, co)
    op()
    got = op.buffer("outData").get(queue)
    ref = orc.beamform(re, co, signed_input=signed)
    scale = np.abs(orc.beamform(re, np.abs(co), signed_input=False))  # sum |x||w|
    assert np.all(np.abs(got - ref) <= 1e-5 * scale + 1e-6)


TC_CASES = [  # B, C, T, A, M, signed: shapes of the tcgen05 stand-alone kernel
    (1, 3, 128, 16, 16, False), (2, 3, 64, 64, 16, False), (1, 5, 256, 64, 64, False), (1, 2, 384, 80, 32, True),
    (1, 2, 256, 8, 2, False), (1, 3, 640, 72, 6, False), (1, 2, 256, 136, 130, False), (1, 2, 48, 24, 70, True),
    (1, 300, 256, 64, 64, False),  # more work items than SMs: the persistent loop wraps and the rings change phase
    # antenna counts whose sample rows are not a multiple of 16 bytes: fetched as eight boxes of an [8 samples x 2A] view
    (1, 3, 256, 197, 256, False), (2, 3, 64, 4, 2, False), (1, 5, 384, 79, 2, True), (1, 2, 48, 23, 4, False),
    (1, 300, 256, 84, 16, False),
    # odd beam counts (8M-byte coefficient / output rows: no tensor map): coefficients by plain loads, register epilogue
    (1, 3, 256, 64, 5, False), (2, 2, 48, 23, 3, True), (1, 2, 384, 80, 97, False), (1, 150, 128, 16, 1, False),
]


@pytest.mark.parametrize("case", TC_CASES, ids=lambda c: "B{}C{}T{}A{}M{}s{:d}".format(*c))
def test_beamform_tensor_core_kernel(dropin, case):
    """dcbf_beamform on its tcgen05 path (bf16 x 3-term bf16 split of ARBITRARY float32 coefficients, here spanning
    2^-20 .. 2^20): within a few float32 ulps of sum |x||w| of the float64 value, like the float32 CUDA-core kernel
    it is cross-checked against (DCBF_FLAG_DEBUG_CUDA_CORES); every output element is written."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, c, t, n_ants, n_beams, signed = case
    rng = np.random.default_rng(31)
    re = rng.integers(0, 256, (b, 2, c, t // 16, 16, n_ants, 2), dtype=np.uint8)
    co = rng.standard_normal((b, 2, c, 2 * n_ants, 2 * n_beams)).astype(np.float32)
    co *= np.exp2(rng.integers(-20, 21, co.shape)).astype(np.float32)
    dev = torch.device("cuda", 0)
    d_re, d_co = torch.from_numpy(re).to(dev), torch.from_numpy(co).to(dev)
    outs = []
    for extra in (0, _capi.FLAG_DEBUG_CUDA_CORES):
        out = torch.full((b, 2, c, t // 16, 16, 2 * n_beams), float("nan"), dtype=torch.float32, device=dev)
        _capi.beamform(d_re, d_co, out, b, c, t, n_ants, n_beams, (_capi.FLAG_SIGNED_INPUT if signed else 0) | extra)
        torch.cuda.synchronize()
        _capi.fused_status()  # the in-kernel watchdog reports through the same status block
        outs.append(out.cpu().numpy())
    ref = orc.beamform(re, co.astype(np.float64), signed_input=signed)
    xs = re.view(np.int8) if signed else re
    scale = np.abs(orc.beamform(np.abs(xs.astype(np.int16)).astype(np.uint8), np.abs(co).astype(np.float64), signed_input=False))
    for got in outs:
        assert not np.isnan(got).any()
        assert np.all(np.abs(got - ref) <= 4e-6 * scale + 1e-30)
    assert np.all(np.abs(outs[0] - outs[1]) <= 4e-6 * scale + 1e-30)


@pytest.mark.parametrize("n_ants", [4, 64, 79, 19])
def test_beamform_coeff_plus_mult_reference_inputs(dropin, n_ants):
    """beamform_mult_kernel_test.py: coefficient op + multiply op, uniform delays, rtol=atol=1e-4."""
    from beamforming.coeff_generator import CoeffGeneratorTemplate
    from beamforming.matrix_multiply import MatrixMultiplyTemplate

    ctx, queue = dropin
    b, t, m, n, xid = 3, 256, 2, 1024, 0
    c = max(1, n // n_ants // 4)
    cg = CoeffGeneratorTemplate(ctx, b, 2, c, n, t // 16, 16, n_ants, m, xid, TS).instantiate(queue)
    mm = MatrixMultiplyTemplate(ctx, n_ants, c, t, m, b).instantiate(queue)
    cg.ensure_all_bound()
    mm.bind(inCoeffs=cg.buffer("outCoeffs"))
    mm.ensure_all_bound()
    rng = np.random.default_rng(seed=2021)
    re = rng.uniform(0, 255, mm.buffer("inData").shape).astype(np.uint8)
    dv = orc.make_delay_vals_uniform(c, m, n_ants)
    cg.buffer("delay_vals").set(queue, dv)
    mm.buffer("inData").set(queue, re)
    cg()
    mm()
    got = mm.buffer("outData").get(queue)
    co = orc.steering_coeffs(dv, b, 2, c, n, n_ants, m, xid, TS)
    ref64 = orc.beamform(re, co)
    ulp = 2.0 ** -23 * orc.beamform_abs_bound(re)[..., None]  # one float32 ulp of the accumulated magnitude
    tol = 1e-4 + 1e-4 * np.abs(ref64) + ulp  # reference tolerance (beamform_mult_kernel_test.py:267) + accumulate ulp
    assert np.all(np.abs(got - ref64) <= tol)
    assert np.all(np.abs(got - orc.complex_mult_beam0(re, co)) <= tol + ulp * np.sqrt(2 * n_ants))


@pytest.mark.parametrize("n_ants", REF_ANTS)
def test_op_sequence_reference_inputs(dropin, n_ants):
    """beamform_op_sequence_test.py protocol, inputs and tolerance (rtol=atol=1e-4), fused kernel."""
    from beamforming.beamform_op_sequence import OpSequenceTemplate

    ctx, queue = dropin
    b, t, m, n, xid = 3, 256, 2, 1024, 0
    c = max(1, n // n_ants // 4)
    op = OpSequenceTemplate(ctx, b, 2, c, n, t // 16, 16, n_ants, m, xid, TS, t).instantiate(queue)
    op.ensure_all_bound()
    buf_in, buf_dv, buf_out = op.buffer("bufin_reorder"), op.buffer("bufin_delay_vals"), op.buffer("bufout_mult")
    host_in = buf_in.empty_like()
    rng = np.random.default_rng(seed=2021)
    host_in[:] = rng.uniform(0, 255, host_in.shape).astype(np.uint8)
    host_dv = buf_dv.empty_like()
    host_dv[:] = orc.make_delay_vals_uniform(c, m, n_ants, samples_delay=5, phase=math.pi / 2)
    buf_in.set(queue, host_in)
    buf_dv.set(queue, host_dv)
    op()
    host_out = buf_out.empty_like()
    buf_out.get(queue, host_out)
    from dpdk_dc_sand_b200 import _capi

    _capi.fused_status()
    x, dv = np.asarray(host_in), np.asarray(host_dv)
    co = orc.steering_coeffs(dv, b, 2, c, n, n_ants, m, xid, TS)
    re = orc.reorder(x)
    # The reference asserts rtol=atol=1e-4 between two float32 evaluations that share their summation order
    # (beamform_op_sequence_test.py:198).  Against the float64 value of the same sum that tolerance needs one
    # float32 ulp of the accumulated magnitude added (sum_a |x_a| ~ 1e3..3e4 here): 2^-23 * sum|x|.
    ref64 = orc.beamform(re, co)
    tol = 1e-4 + 1e-4 * np.abs(ref64) + 2.0 ** -23 * _budget(x) * 2.0 ** 10
    assert np.all(np.abs(host_out - ref64) <= tol)
    chk = orc.complex_mult_beam0(re, co)  # the reference's own float32 checker (beam-uniform coefficients here)
    assert np.all(np.abs(host_out - chk) <= tol + 2.0 ** -23 * _budget(x) * 2.0 ** 10 * np.sqrt(2 * n_ants))
    assert np.all(np.abs(host_out - orc.beamform_pipeline(x, dv, n, xid, TS)) <= _budget(x))


FUSED_CASES = [
    # B, A, C, T, M, N, xeng_id, signed, fp16_coeff
    (1, 600, 2, 128, 8, 64, 0, False, False),     # more antennas than any whole B tile set holds: k-block ring only
    (1, 600, 2, 400, 8, 64, 1, True, False),      # ... and more than two time tiles: the ring is re-streamed per pair
    (1, 4, 64, 256, 4, 64, 0, False, False),      # BASELINE configs[0]
    (1, 64, 24, 256, 16, 1024, 0, False, False),  # configs[1] geometry, a slice of channels
    (1, 64, 10, 256, 64, 4096, 5, False, False),  # configs[2] geometry, xeng 5 of 8... (C=10 slice)
    (1, 80, 6, 256, 32, 32768, 3, False, False),  # configs[3] geometry
    (1, 197, 3, 256, 256, 4096, 1, False, False),  # configs[4] geometry (16 N tiles, 7 k-blocks)
    (2, 5, 3, 32, 3, 256, 1, False, False),       # odd everything, partial time tile
    (3, 23, 7, 48, 2, 1024, 0, True, False),      # signed input
    (1, 64, 10, 256, 64, 4096, 0, False, True),   # single-fp16 coefficients (fast mode)
    (1, 33, 200, 144, 9, 4096, 2, False, False),  # more channels than SMs, T = 128 + 16
    (1, 256, 2, 256, 2, 1024, 0, False, False),   # 8 k-blocks
    (1, 300, 2, 64, 5, 512, 0, False, False),     # more antennas than coefficient threads, odd beams, 16-column N tiles
    (2, 16, 3, 128, 130, 256, 1, False, False),   # 260 columns: three N tiles of 96 (TMA boxes inside their tile)
    (1, 8, 1, 16, 1, 8, 0, False, False),         # smallest legal problem
    (1, 64, 149, 16, 4, 4096, 3, True, False),    # one more channel than SMs, 16-sample heaps, signed
]


@pytest.mark.parametrize("case", FUSED_CASES, ids=lambda c: "B{}A{}C{}T{}M{}N{}x{}s{:d}h{:d}".format(*c))
def test_fused_random_delays(dropin, case):
    from beamforming.beamform_op_sequence import OpSequenceTemplate
    from dpdk_dc_sand_b200 import _capi

    ctx, queue = dropin
    b, a, c, t, m, n, xid, signed, fp16 = case
    op = OpSequenceTemplate(ctx, b, 2, c, n, t // 16, 16, a, m, xid, TS, t).instantiate(queue)
    op.signed_input, op.fp16_coeff = signed, fp16
    op.ensure_all_bound()
    assert not op.slots["bufint_coeff"].is_bound and not op.slots["bufint_data"].is_bound
    x = orc.make_samples(b, a, c, t, seed=100 + a)
    dv = orc.make_delay_vals_random(c, m, a, seed=200 + m)
    op.buffer("bufin_reorder").set(queue, x)
    op.buffer("bufin_delay_vals").set(queue, dv)
    op.buffer("bufout_mult").zero(queue)
    n0 = _capi.launch_count()
    op()
    got = op.buffer("bufout_mult").get(queue).astype(np.float64)
    _capi.fused_status()
    assert _capi.launch_count() - n0 == 1
    ref = orc.beamform_pipeline(x, dv, n, xid, TS, signed_input=signed)
    err = np.abs(got - ref)
    assert not np.isnan(got).any()
    assert np.all(err <= _budget(x, signed)), f"max err {err.max()} vs budget"
    if not fp16:  # hi+lo coefficients: float32-grade result
        assert np.all(err <= 2.0 ** -18 * _budget(x, signed) * 2 ** 10 + 1e-3)


@pytest.mark.parametrize("max_delay_samples", [1e5, 2e7], ids=["58us", "12ms_float64_path"])
def test_fused_large_delays(dropin, max_delay_samples):
    """Geometric-scale delays (tens of microseconds: ~1e4..1e5 half-turns of phase at the band edge) and an
    absurd 12 ms that forces the kernel's float64 phase path: the float-pair phase arithmetic must still
    agree with the float64 oracle to float32-grade accuracy."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m, n, xid = 1, 64, 6, 128, 8, 4096, 7  # band edge: channels 3584..3589 of 4096
    x = orc.make_samples(b, a, c, t, seed=31)
    dv = orc.make_delay_vals_random(c, m, a, seed=32, max_delay_samples=max_delay_samples)
    out = torch.empty((b, 2, c, t // 16, 16, 2 * m), dtype=torch.float32, device="cuda")
    _capi.fused(torch.from_numpy(x).cuda(), torch.from_numpy(dv).cuda(), out, b, a, c, n, t, m, xid, TS)
    _capi.fused_status()
    ref = orc.beamform_pipeline(x, dv, n, xid, TS)
    err = np.abs(out.cpu().numpy().astype(np.float64) - ref)
    assert np.all(err <= 2.0 ** -8 * _budget(x) + 1e-3), f"max err {err.max()}"  # 2^-18 * sum|x|


def test_host_reorder_helper_runs_on_the_gpu(dropin):
    """``beamforming.reorder.reorder`` (reference: beamforming/reorder.py:46) keeps its signature but has no CPU
    implementation here: it must go through the CUDA kernel and agree with the oracle bit for bit."""
    from beamforming import reorder as host_reorder
    from dpdk_dc_sand_b200 import _capi

    x = orc.make_samples(2, 9, 5, 48, seed=77)
    n0 = _capi.launch_count()
    got = host_reorder.reorder(x, x.shape, (2, 2, 5, 3, 16, 9, 2))
    assert _capi.launch_count() == n0 + 1
    np.testing.assert_array_equal(got, orc.reorder(x))
    with pytest.raises(ValueError):
        host_reorder.reorder(x, x.shape, (2, 2, 5, 48, 1, 9, 2))


def _tv_delay_vals(c, m, a, seed):
    """Realistic rate magnitudes: delay rate +-2e-9 s/s, phase rate +-2 rad/s (random junk rates would swamp the phase)."""
    rng = np.random.default_rng(seed)
    dv = orc.make_delay_vals_random(c, m, a, seed=seed, max_delay_samples=2000.0)
    dv[..., 1] = rng.uniform(-2e-9, 2e-9, dv.shape[:3]).astype(np.float32)
    dv[..., 3] = rng.uniform(-2.0, 2.0, dv.shape[:3]).astype(np.float32)
    return dv


def test_time_varying_steering(dropin):
    """Next-row feature: per-heap delay/phase rates.  Stand-alone coefficients <= 1e-6 of the float64
    oracle, fused beams inside the budget at float32 grade, the fused and three-kernel paths agree, and
    batch_times = 0 reproduces the static path (to float32 rounding: the static path uses a leaner phase evaluation)."""
    from beamforming.beamform_op_sequence import OpSequenceTemplate

    ctx, queue = dropin
    b, a, c, t, m, n, xid = 3, 64, 5, 128, 6, 4096, 6
    times = [0.0, 1.37, -7.25]
    x = orc.make_samples(b, a, c, t, seed=41)
    dv = _tv_delay_vals(c, m, a, seed=42)
    ref = orc.beamform_pipeline(x, dv, n, xid, TS, batch_dt=times)
    assert np.abs(ref[1] - orc.beamform_pipeline(x, dv, n, xid, TS)[1]).max() > 100.0  # the rates matter
    outs = {}
    for fused in (True, False):
        op = OpSequenceTemplate(ctx, b, 2, c, n, t // 16, 16, a, m, xid, TS, t).instantiate(queue)
        op.fused, op.materialize_intermediates, op.batch_times = fused, True, times
        op.ensure_all_bound()
        op.buffer("bufin_reorder").set(queue, x)
        op.buffer("bufin_delay_vals").set(queue, dv)
        op()
        outs[fused] = op.buffer("bufout_mult").get(queue).astype(np.float64)
        co = op.buffer("bufint_coeff").get(queue).astype(np.float64)
        ref_co = orc.steering_coeffs(dv, b, 2, c, n, a, m, xid, TS, out_dtype=np.float64, batch_dt=times)
        assert np.abs(co - ref_co).max() <= 1e-6
    err = np.abs(outs[True] - ref)
    assert np.all(err <= 2.0 ** -8 * _budget(x) + 1e-3), f"max err {err.max()}"
    np.testing.assert_allclose(outs[True], outs[False], rtol=0, atol=float(_budget(x).max()) * 2.0 ** -6)
    # dt = 0 for every heap == the static kernel, to the rounding of the two phase evaluations (<= 3e-7 rad)
    res = []
    for bt in (None, [0.0] * b):
        op = OpSequenceTemplate(ctx, b, 2, c, n, t // 16, 16, a, m, xid, TS, t).instantiate(queue)
        op.batch_times = bt
        op.ensure_all_bound()
        op.buffer("bufin_reorder").set(queue, x)
        op.buffer("bufin_delay_vals").set(queue, dv)
        op()
        res.append(op.buffer("bufout_mult").get(queue))
    assert np.all(np.abs(res[0].astype(np.float64) - res[1]) <= 2.0 ** -10 * _budget(x))  # 2^-20 * sum|x|


def test_time_varying_steering_with_many_antennas_and_beams(dropin):
    """Per-heap times in the K-streamed mode (a shape whose B tile sets need several N tiles, heaps of three time
    tiles): inside the budget against the float64 oracle and equal, to float32 rounding, to the whole-tile-set mode."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m, n, xid = 3, 100, 3, 384, 100, 512, 1
    assert _capi.fused_tiling(a, m)[2] > 1
    times = [0.0, 2.5, -3.75]
    x = orc.make_samples(b, a, c, t, seed=43)
    dv = _tv_delay_vals(c, m, a, seed=44)
    ref = orc.beamform_pipeline(x, dv, n, xid, TS, batch_dt=times)
    dx, ddv = torch.from_numpy(x).cuda(), torch.from_numpy(dv).cuda()
    outs = []
    for flags in (0, _capi.FLAG_DEBUG_NO_KSTREAM):
        out = torch.full(ref.shape, float("nan"), dtype=torch.float32, device="cuda")
        _capi.fused(dx, ddv, out, b, a, c, n, t, m, xid, TS, flags, batch_dt=times)
        torch.cuda.synchronize()
        _capi.fused_status()
        outs.append(out.cpu().numpy().astype(np.float64))
        assert np.all(np.abs(outs[-1] - ref) <= 2.0 ** -8 * _budget(x) + 1e-3)
    np.testing.assert_allclose(outs[0], outs[1], rtol=0, atol=float(_budget(x).max()) * 2.0 ** -6)


@pytest.mark.parametrize("case", [(1, 2, 5, 64, 64, 4096, 3), (2, 1, 3, 33, 20, 256, 0), (1, 2, 2, 80, 100, 1024, 1),
                                  (3, 2, 4, 7, 6, 64, 0), (1, 1, 2, 16, 3, 64, 0)],
                         ids=["M64_whole_rows", "M20", "M100_two_beam_tiles", "M6_plain_stores_for_f16", "M3_plain_stores"])
def test_coeff_generator_bulk_tiles_and_half_output(dropin, case):
    """The stand-alone coefficient kernel stages every tile in the output layout and writes the (batch, pol) replicas
    with bulk copies; odd row pitches take plain stores.  float32: equal to the float64 oracle to 1e-6 in every
    replica.  float16 (the precursor's 16-bit output, BeamformerKernels.cu:172-185): exactly the round-to-nearest fp16
    of the float32 output."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, p, c, a, m, n, xid = case
    dv = orc.make_delay_vals_random(c, m, a, seed=81)
    ref = orc.steering_coeffs(dv, b, p, c, n, a, m, xid, TS, out_dtype=np.float64)
    ddv = torch.from_numpy(dv).cuda()
    out32 = torch.full(ref.shape, float("nan"), dtype=torch.float32, device="cuda")
    out16 = torch.full(ref.shape, float("nan"), dtype=torch.float16, device="cuda")
    _capi.coeffs(ddv, out32, b, p, c, n, a, m, xid, TS)
    _capi.coeffs(ddv, out16, b, p, c, n, a, m, xid, TS)
    torch.cuda.synchronize()
    got32, got16 = out32.cpu().numpy(), out16.cpu().numpy()
    assert np.abs(got32.astype(np.float64) - ref).max() <= 1e-6
    np.testing.assert_array_equal(got16, got32.astype(np.float16))


@pytest.mark.parametrize("case", [(1, 197, 4, 256, 256), (1, 100, 5, 384, 100), (2, 197, 3, 256, 130), (1, 520, 3, 160, 70),
                                  (1, 64, 9, 256, 200), (3, 33, 4, 640, 97)],
                         ids=["c5_like", "three_time_tiles", "two_heaps_ragged_n", "520_antennas_ragged_t", "64_antennas",
                              "odd_beams_five_tiles"])
def test_cta_pair_mode(dropin, case):
    """Many antennas x beams run on CTA pairs (cta_group::2 MMAs of M = 256, two time tiles and half of the coefficients
    per CTA): inside the budget against the float64 oracle, and bit-identical to the single-CTA K-streamed kernel
    (same products, same accumulation order), including heaps whose last pair has one tile, ragged last tiles and N tiles."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m = case
    n, xid = 1024, 2
    x = orc.make_samples(b, a, c, t, seed=71)
    dv = orc.make_delay_vals_random(c, m, a, seed=72)
    ref = orc.beamform_pipeline(x, dv, n, xid, TS)
    dx, ddv = torch.from_numpy(x).cuda(), torch.from_numpy(dv).cuda()
    outs = []
    for flags in (0, _capi.FLAG_DEBUG_NO_PAIR):
        out = torch.full(ref.shape, float("nan"), dtype=torch.float32, device="cuda")
        _capi.fused(dx, ddv, out, b, a, c, n, t, m, xid, TS, flags)
        torch.cuda.synchronize()
        _capi.fused_status()
        outs.append(out.cpu().numpy())
    assert np.all(np.abs(outs[0].astype(np.float64) - ref) <= _budget(x))
    np.testing.assert_array_equal(outs[0], outs[1])


@pytest.mark.parametrize("case", [(1, 64, 10, 256, 64), (2, 64, 160, 128, 64), (1, 32, 13, 384, 128), (1, 23, 50, 256, 60),
                                  (1, 64, 75, 256, 40)],
                         ids=["ten_channels_all_cut", "one_round_plus_12_channels", "256_column_tile", "ragged_second_half",
                              "ineligible_tile_list_only"])
def test_last_round_channels_cut_into_beam_halves(dropin, case):
    """The channels of the last scheduling round are cut into pieces so that all CTAs stay busy: halves of the N tile's
    beams first (64- or 128-column MMAs, half the coefficient work per piece), then along the tile list.  Inside the budget
    against the float64 oracle and bit-identical to the uncut schedule and to the tile-list-only cut."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m = case
    n, xid = 1024, 1
    x = orc.make_samples(b, a, c, t, seed=81)
    dv = orc.make_delay_vals_random(c, m, a, seed=82)
    ref = orc.beamform_pipeline(x, dv, n, xid, TS)
    dx, ddv = torch.from_numpy(x).cuda(), torch.from_numpy(dv).cuda()
    outs = []
    for flags in (0, _capi.FLAG_DEBUG_NO_BEAM_PIECES, _capi.FLAG_DEBUG_WHOLE_CHANNELS):
        out = torch.full(ref.shape, float("nan"), dtype=torch.float32, device="cuda")
        _capi.fused(dx, ddv, out, b, a, c, n, t, m, xid, TS, flags)
        torch.cuda.synchronize()
        _capi.fused_status()
        outs.append(out.cpu().numpy())
    assert np.all(np.abs(outs[0].astype(np.float64) - ref) <= _budget(x))
    for other in outs[1:]:
        np.testing.assert_array_equal(outs[0], other)


@pytest.mark.parametrize("case", [(2, 64, 3, 512, 6, 4096, 6), (1, 100, 2, 384, 100, 512, 1), (2, 16, 5, 208, 4, 256, 0)],
                         ids=["whole_tile_sets", "k_streamed", "ragged_last_tile"])
def test_sub_heap_time_varying_steering(dropin, case):
    """Steering that follows the delay model INSIDE a heap (native precursor: coefficients per timestamp,
    BeamformerKernels.cu:153-167): with sample_dt every 128-sample time tile has its own coefficient set, evaluated at
    the tile's centre.  (1) the kernel equals the oracle's restatement of exactly that; (2) against the exact
    per-sample steering the result is inside the 2^-10 * sum|x| budget, where one set per heap is not."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m, n, xid = case
    sample_dt = 8192 / 1712e6  # the precursor's FFT_SIZE * SAMPLING_PERIOD at the MeerKAT L-band sample rate
    times = [0.4 + 0.01 * i for i in range(b)]
    x = orc.make_samples(b, a, c, t, seed=45)
    dv = _tv_delay_vals(c, m, a, seed=46)
    exact = orc.beamform_pipeline(x, dv, n, xid, TS, batch_dt=times, sample_dt=sample_dt)
    per_tile = orc.beamform_pipeline(x, dv, n, xid, TS, batch_dt=times, sample_dt=sample_dt, tile=128)
    dx, ddv = torch.from_numpy(x).cuda(), torch.from_numpy(dv).cuda()
    got = {}
    for dt in (sample_dt, 0.0):
        out = torch.full(exact.shape, float("nan"), dtype=torch.float32, device="cuda")
        _capi.fused_ex(dx, ddv, out, b, a, c, n, t, m, xid, TS, batch_dt=times, sample_dt=dt)
        torch.cuda.synchronize()
        _capi.fused_status()
        got[dt] = out.cpu().numpy().astype(np.float64)
    budget = _budget(x)
    assert np.all(np.abs(got[sample_dt] - per_tile) <= 2.0 ** -8 * budget + 1e-3)   # the kernel does what it says
    assert np.all(np.abs(got[sample_dt] - exact) <= budget)                          # and that is inside the budget
    if t >= 384:  # one set per heap (at the heap's first sample) is not: the drift over >= 384 samples is too large
        assert np.any(np.abs(got[0.0] - exact) > budget)


GUARD_CASES = [(1, 64, 5, 256, 64), (2, 23, 3, 48, 3), (1, 197, 3, 256, 256), (1, 100, 2, 384, 100), (1, 520, 2, 160, 70),
               (1, 64, 150, 128, 64), (3, 33, 4, 640, 97), (1, 1, 7, 16, 1), (1, 80, 9, 272, 32)]


@pytest.mark.parametrize("case", GUARD_CASES, ids=lambda c: "B{}A{}C{}T{}M{}".format(*c))
def test_outputs_stay_inside_their_buffers(dropin, case):
    """Bounds check of our own (no memory checker on the GPU boxes): every output of the C-ABI lives in the middle of a
    larger allocation whose 64 KiB on either side hold a canary pattern; after the fused kernel (float32 and int8, every
    tiling mode the shapes select: whole tile sets, pieces, K-streamed, CTA pairs), the stand-alone reorder / coefficients /
    contraction (tcgen05 and CUDA cores) the canaries are intact and the inputs are unchanged."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m = case
    n, xid, guard = 4 * c, 1, 1 << 16
    g = torch.Generator(device="cuda").manual_seed(5)

    def guarded(shape, dtype):
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        raw = torch.full((guard + nbytes + guard,), 0xA5, dtype=torch.uint8, device="cuda")
        return raw, raw[guard:guard + nbytes].view(dtype).view(shape)

    def intact(raw):
        return bool((raw[:guard] == 0xA5).all()) and bool((raw[-guard:] == 0xA5).all())

    x = torch.randint(0, 256, (b, a, c, t, 2, 2), dtype=torch.uint8, device="cuda", generator=g)
    dv = torch.rand((c, m, a, 4), dtype=torch.float32, device="cuda", generator=g) * 2e-7
    gains = torch.full((m,), 0.01, dtype=torch.float32, device="cuda")
    x0, dv0 = x.clone(), dv.clone()
    k = t // 16
    raws = {}
    raws["fused"], out = guarded((b, 2, c, k, 16, 2 * m), torch.float32)
    _capi.fused(x, dv, out, b, a, c, n, t, m, xid, TS)
    raws["fused_q8"], out8 = guarded((b, 2, c, k, 16, 2 * m), torch.int8)
    _capi.fused_q8(x, dv, gains, out8, b, a, c, n, t, m, xid, TS)
    raws["reorder"], re = guarded((b, 2, c, k, 16, a, 2), torch.uint8)
    _capi.reorder(x, re, b, a, c, t)
    raws["coeffs"], co = guarded((b, 2, c, 2 * a, 2 * m), torch.float32)
    _capi.coeffs(dv, co, b, 2, c, n, a, m, xid, TS)
    raws["coeffs_f16"], co16 = guarded((b, 2, c, 2 * a, 2 * m), torch.float16)
    _capi.coeffs(dv, co16, b, 2, c, n, a, m, xid, TS)
    raws["beamform"], bf = guarded((b, 2, c, k, 16, 2 * m), torch.float32)
    _capi.beamform(re, co, bf, b, c, t, a, m)
    raws["beamform_cuda_cores"], bf2 = guarded((b, 2, c, k, 16, 2 * m), torch.float32)
    _capi.beamform(re, co, bf2, b, c, t, a, m, _capi.FLAG_DEBUG_CUDA_CORES)
    torch.cuda.synchronize()
    _capi.fused_status()
    for name, raw in raws.items():
        assert intact(raw), name
    assert torch.equal(x, x0) and torch.equal(dv, dv0)
    assert not bool(torch.isnan(out).any()) and not bool(torch.isnan(bf).any())
    budget = x.to(torch.float32).abs().sum(dim=(1, 5)).amax() * 2.0 ** -10
    assert float((out - bf).abs().max()) <= 2 * float(budget)  # the two paths agree as well


@pytest.mark.parametrize("case", [(1, 64, 7, 256, 16, 1024, 0, False), (2, 23, 3, 48, 3, 256, 1, True),
                                  (1, 80, 4, 256, 32, 32768, 3, False), (1, 4, 9, 128, 8, 64, 0, False),
                                  (1, 1, 21, 128, 15, 21, 0, False), (2, 100, 3, 384, 100, 512, 1, False),
                                  (1, 600, 2, 128, 5, 64, 0, True)],
                         ids=["M16_tma", "M3_ragged_signed", "M32_A80", "M8_16cols", "single_antenna", "k_streamed",
                              "k_streamed_600_antennas_odd_beams"])
def test_fused_q8_requantised_output(dropin, case):
    """Next-row feature: int8 beams = clip(rint(beam * gain[m]), -127, 127) in the fused epilogue.
    Against the float64 oracle: never more than one quantisation step away, practically always equal (the only
    differences are float32-vs-float64 ties at .5 boundaries), same saturation count."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m, n, xid, signed = case
    x = orc.make_samples(b, a, c, t, seed=51)
    dv = orc.make_delay_vals_random(c, m, a, seed=52)
    ref = orc.beamform_pipeline(x, dv, n, xid, TS, signed_input=signed)
    rng = np.random.default_rng(53)
    gains = (rng.uniform(0.5, 3.0, m) * 100.0 / np.abs(ref).max()).astype(np.float32)  # a few percent saturate
    want, want_clipped = orc.requantise(ref, gains)
    out = torch.full(ref.shape, 77, dtype=torch.int8, device="cuda")
    sat = torch.zeros(1, dtype=torch.int64, device="cuda")
    flags = _capi.FLAG_SIGNED_INPUT if signed else 0
    _capi.fused_q8(torch.from_numpy(x).cuda(), torch.from_numpy(dv).cuda(), torch.from_numpy(gains).cuda(), out,
                   b, a, c, n, t, m, xid, TS, flags=flags, saturated=sat)
    _capi.fused_status()
    got = out.cpu().numpy()
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert diff.max() <= 1
    assert np.count_nonzero(diff) <= 1e-3 * diff.size
    assert want_clipped > 0 and abs(int(sat.item()) - want_clipped) <= 2 + 1e-3 * want_clipped
    assert _capi.fused_q8_bytes(b, a, c, t, m) == x.size + dv.size * 4 + got.size + 4 * m
    # without the saturation counter the epilogue clamps pairs of 16-bit results (packed add-min-relu) whenever
    # 510 * A * max|gain| < 2^15 steps, which these gains satisfy: the bytes must not change
    assert 510.0 * a * float(gains.max()) < 32000.0
    out2 = torch.full(ref.shape, 77, dtype=torch.int8, device="cuda")
    _capi.fused_q8(torch.from_numpy(x).cuda(), torch.from_numpy(dv).cuda(), torch.from_numpy(gains).cuda(), out2,
                   b, a, c, n, t, m, xid, TS, flags=flags)
    _capi.fused_status()
    np.testing.assert_array_equal(out2.cpu().numpy(), got)


def test_q8_operator_and_host_plan(dropin):
    """The q8 extension through the operator API and through the host-buffer C-ABI plan give the same bytes."""
    from beamforming.beamform_op_sequence import QuantisedOpSequenceTemplate
    from dpdk_dc_sand_b200 import _capi

    ctx, queue = dropin
    b, a, c, t, m, n, xid = 2, 16, 21, 64, 8, 256, 1
    x = orc.make_samples(b, a, c, t, seed=61)
    dv = orc.make_delay_vals_random(c, m, a, seed=62)
    ref = orc.beamform_pipeline(x, dv, n, xid, TS)
    gains = np.full(m, 90.0 / np.abs(ref).max(), np.float32)
    op = QuantisedOpSequenceTemplate(ctx, b, 2, c, n, t // 16, 16, a, m, xid, TS, t).instantiate(queue)
    op.ensure_all_bound()
    op.buffer("bufin_reorder").set(queue, x)
    op.buffer("bufin_delay_vals").set(queue, dv)
    op.buffer("bufin_gains").set(queue, gains)
    op()
    got = op.buffer("bufout_q8").get(queue)
    assert got.dtype == np.int8 and got.shape == ref.shape and op.saturated == 0
    want, _ = orc.requantise(ref, gains)
    assert np.abs(got.astype(np.int32) - want).max() <= 1 and np.count_nonzero(got != want) <= 1e-3 * want.size
    plan = _capi.HostPlan(b, a, c, n, t, m, xid, TS, chunk_chans=8, n_slots=2)
    plan.set_gains(gains)
    host = np.zeros(ref.shape, np.int8)
    assert plan.run_q8(x, dv, host) == 0
    plan.close()
    np.testing.assert_array_equal(host, got)


def test_beam_weights(dropin):
    """Next-row feature: per-(beam, antenna) real weights (the ?beam-weights payload) folded into the
    coefficients, alone and together with per-heap times; fused == three-kernel chain == float64 oracle."""
    from beamforming.beamform_op_sequence import OpSequenceTemplate

    ctx, queue = dropin
    b, a, c, t, m, n, xid = 2, 33, 6, 64, 5, 1024, 2
    x = orc.make_samples(b, a, c, t, seed=71)
    dv = _tv_delay_vals(c, m, a, seed=72)
    w = np.random.default_rng(73).uniform(0.0, 2.0, (m, a)).astype(np.float32)
    w[1, :] = 0.0  # a muted beam
    for times in (None, [0.0, 2.5]):
        ref = orc.beamform_pipeline(x, dv, n, xid, TS, batch_dt=times, weights=w)
        outs = {}
        for fused in (True, False):
            op = OpSequenceTemplate(ctx, b, 2, c, n, t // 16, 16, a, m, xid, TS, t).instantiate(queue)
            op.fused, op.materialize_intermediates, op.batch_times, op.beam_weights = fused, True, times, w
            op.ensure_all_bound()
            op.buffer("bufin_reorder").set(queue, x)
            op.buffer("bufin_delay_vals").set(queue, dv)
            op()
            outs[fused] = op.buffer("bufout_mult").get(queue).astype(np.float64)
            co = op.buffer("bufint_coeff").get(queue).astype(np.float64)
            ref_co = orc.steering_coeffs(dv, b, 2, c, n, a, m, xid, TS, out_dtype=np.float64,
                                         batch_dt=times if times else None, weights=w)
            assert np.abs(co - ref_co).max() <= 2e-6  # weights up to 2
        err = np.abs(outs[True] - ref)
        assert np.all(err <= 2.0 * (2.0 ** -8 * _budget(x) + 1e-3)), f"max err {err.max()}"
        assert np.abs(outs[True][:, :, :, :, :, 2:4]).max() == 0.0  # beam 1 is muted exactly
        np.testing.assert_allclose(outs[True], outs[False], rtol=0, atol=float(_budget(x).max()) * 2.0 ** -5)


@pytest.mark.parametrize("scale", [1e-4, 1.0, 37.0, 1e4], ids=["w_1e-4", "w_1", "w_37", "w_1e4"])
def test_beam_weights_of_any_magnitude(dropin, scale):
    """Weights far from 1 (|w| > 64 used to overflow the fp16 coefficient pair, |w| < 1e-3 lost its
    residual to fp16 subnormals): the operator passes a power-of-two bound of the weights (beam_weights_log2), the
    kernel moves it from the coefficients onto the voltages' conversion exponent, and the result keeps the same
    RELATIVE accuracy at every magnitude -- inside the budget scaled by max|w|, and equal to the float64-weighted
    three-kernel chain."""
    from beamforming.beamform_op_sequence import OpSequenceTemplate

    ctx, queue = dropin
    b, a, c, t, m, n, xid = 1, 40, 5, 128, 6, 1024, 1
    x = orc.make_samples(b, a, c, t, seed=75)
    dv = orc.make_delay_vals_random(c, m, a, seed=76)
    w = (np.random.default_rng(77).uniform(0.1, 1.0, (m, a)) * scale).astype(np.float32)
    ref = orc.beamform_pipeline(x, dv, n, xid, TS, weights=w)
    outs = {}
    for fused in (True, False):
        op = OpSequenceTemplate(ctx, b, 2, c, n, t // 16, 16, a, m, xid, TS, t).instantiate(queue)
        op.fused, op.beam_weights = fused, w
        op.ensure_all_bound()
        op.buffer("bufin_reorder").set(queue, x)
        op.buffer("bufin_delay_vals").set(queue, dv)
        op()
        outs[fused] = op.buffer("bufout_mult").get(queue).astype(np.float64)
        assert np.all(np.isfinite(outs[fused]))
    wmax = float(np.abs(w).max())
    assert np.all(np.abs(outs[True] - ref) <= 2.0 ** -10 * wmax * _budget(x))  # 2^-20 * max|w| * sum|x|
    np.testing.assert_allclose(outs[True], outs[False], rtol=0, atol=float(_budget(x).max()) * wmax * 2.0 ** -8)


def test_delay_model_updater_switches_at_a_heap_boundary(dropin):
    """Double-buffered delay_vals / weights update: heaps launched before activate() use the old
    model, heaps launched after it use the new one, and the upload runs on a side stream."""
    from beamforming.beamform_op_sequence import OpSequenceTemplate
    from dpdk_dc_sand_b200.delay_model import DelayModelUpdater

    ctx, queue = dropin
    b, a, c, t, m, n, xid = 1, 16, 9, 64, 4, 64, 0
    x = orc.make_samples(b, a, c, t, seed=81)
    dv0 = orc.make_delay_vals_random(c, m, a, seed=82)
    dv1 = orc.make_delay_vals_random(c, m, a, seed=83)
    w1 = np.random.default_rng(84).uniform(0.5, 1.5, (m, a)).astype(np.float32)
    op = OpSequenceTemplate(ctx, b, 2, c, n, t // 16, 16, a, m, xid, TS, t).instantiate(queue)
    op.ensure_all_bound()
    op.buffer("bufin_reorder").set(queue, x)
    op.buffer("bufin_delay_vals").set(queue, dv0)
    upd = DelayModelUpdater(op)
    assert not upd.activate()
    upd.update(dv1, w1)  # uploads while ...
    op()                 # ... this heap still sees model 0
    out0 = op.buffer("bufout_mult").get(queue).astype(np.float64)
    assert upd.activate()
    op()
    out1 = op.buffer("bufout_mult").get(queue).astype(np.float64)
    assert np.all(np.abs(out0 - orc.beamform_pipeline(x, dv0, n, xid, TS)) <= _budget(x))
    assert np.all(np.abs(out1 - orc.beamform_pipeline(x, dv1, n, xid, TS, weights=w1)) <= 1.5 * _budget(x))
    upd.update(dv0)      # back again through the other buffer, weights stay
    assert upd.activate()
    op()
    out2 = op.buffer("bufout_mult").get(queue).astype(np.float64)
    assert np.all(np.abs(out2 - orc.beamform_pipeline(x, dv0, n, xid, TS, weights=w1)) <= 1.5 * _budget(x))


def test_delay_model_updater_repacks_the_coefficients(dropin):
    """The same hook on an operation that runs on packed steering coefficients: activate() packs the new model on the
    compute stream, heaps queued before it still get the old tile sets."""
    from beamforming.beamform_op_sequence import OpSequenceTemplate
    from dpdk_dc_sand_b200 import _capi
    from dpdk_dc_sand_b200.delay_model import DelayModelUpdater

    ctx, queue = dropin
    b, a, c, t, m, n, xid = 1, 64, 30, 256, 16, 1024, 0
    x = orc.make_samples(b, a, c, t, seed=181)
    dv0 = orc.make_delay_vals_random(c, m, a, seed=182)
    dv1 = orc.make_delay_vals_random(c, m, a, seed=183)
    op = OpSequenceTemplate(ctx, b, 2, c, n, t // 16, 16, a, m, xid, TS, t).instantiate(queue)
    op.ensure_all_bound()
    op.buffer("bufin_reorder").set(queue, x)
    op.buffer("bufin_delay_vals").set(queue, dv0)
    assert op.pack_coefficients()
    upd = DelayModelUpdater(op)
    upd.update(dv1)
    op()
    out0 = op.buffer("bufout_mult").get(queue).astype(np.float64)
    first_pack = op._packed[0]
    assert upd.activate()
    assert op._packed is not None and op._packed[0] is not first_pack
    n0 = _capi.launch_count()
    op()
    assert _capi.launch_count() - n0 == 1
    out1 = op.buffer("bufout_mult").get(queue).astype(np.float64)
    _capi.fused_status()
    assert np.all(np.abs(out0 - orc.beamform_pipeline(x, dv0, n, xid, TS)) <= _budget(x))
    assert np.all(np.abs(out1 - orc.beamform_pipeline(x, dv1, n, xid, TS)) <= _budget(x))


def test_ingest_ring_feeds_the_host_plan(dropin):
    """SURVEY 8f-3 end to end: heaps -> page-locked chunk (dcbf_ingest_*) -> dcbf_host_plan_run -> beams."""
    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m, n, xid, step = 2, 6, 10, 32, 3, 64, 0, 4096
    dv = orc.make_delay_vals_random(c, m, a, seed=92)
    ing = _capi.Ingest(4, b, a, c, t, step, pinned=True)
    plan = _capi.HostPlan(b, a, c, n, t, m, xid, TS, chunk_chans=4, n_slots=2)
    rng = np.random.default_rng(93)
    heaps = rng.integers(0, 256, (2 * b, a, c, t, 2, 2), dtype=np.uint8)
    order = [(h, ant) for h in range(2 * b) for ant in range(a)]
    rng.shuffle(order)
    for h, ant in order:
        assert ing.heap(h * step, ant, heaps[h, ant])
    for k in range(2):
        samples, ts, present = ing.pop()
        assert ts == k * b * step and present.all()
        out = np.zeros((b, 2, c, t // 16, 16, 2 * m), np.float32)
        plan.run(samples, dv, out)
        ref = orc.beamform_pipeline(heaps[k * b:(k + 1) * b], dv, n, xid, TS)
        assert np.all(np.abs(out - ref) <= _budget(heaps[k * b:(k + 1) * b]))
        ing.release(samples)
    assert ing.pop() is None
    plan.close()
    ing.close()


def test_concurrent_launches_on_several_streams_do_not_share_the_channel_queue(dropin):
    """The dynamic channel scheduler draws from a per-launch counter slot: launches in flight at the same time on
    different streams (and back-to-back DCBF_FLAG_STREAMING launches on one stream) must each cover every channel
    exactly once -- checked by bit-equality with a serial launch of the same problem."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m, n = 1, 32, 700, 64, 8, 1024
    g = torch.Generator(device="cuda").manual_seed(7)
    xs = [torch.randint(0, 256, (b, a, c, t, 2, 2), dtype=torch.uint8, device="cuda", generator=g) for _ in range(3)]
    dv = torch.from_numpy(orc.make_delay_vals_random(c, m, a, seed=4)).cuda()
    want = []
    for x in xs:
        o = torch.empty((b, 2, c, t // 16, 16, 2 * m), dtype=torch.float32, device="cuda")
        _capi.fused(x, dv, o, b, a, c, n, t, m, 0, TS)
        torch.cuda.synchronize()
        want.append(o)
    streams = [torch.cuda.Stream() for _ in range(3)]
    outs = [[torch.full_like(want[0], float("nan")) for _ in range(4)] for _ in range(3)]
    torch.cuda.synchronize()  # (the fills ran on the default stream, the launches go to side streams)
    for rep in range(4):
        for i, st in enumerate(streams):
            flags = _capi.FLAG_STREAMING if rep else 0
            _capi.fused(xs[i], dv, outs[i][rep], b, a, c, n, t, m, 0, TS, flags, st)
    torch.cuda.synchronize()
    _capi.fused_status()
    for i in range(3):
        for rep in range(4):
            assert torch.equal(outs[i][rep], want[i]), (i, rep)


@pytest.mark.parametrize("case", [(1, 197, 3, 256, 256, 4096, 1), (2, 45, 4, 128, 131, 512, 0), (3, 64, 2, 48, 80, 256, 2),
                                  (1, 33, 150, 16, 70, 4096, 5), (2, 45, 3, 640, 131, 512, 1)],
                         ids=["C5_geometry", "odd_beams_direct_epilogue_2_heaps", "3_heaps_partial_tile", "150_channels",
                              "five_time_tiles"])
def test_k_streamed_b_tiles_match_whole_tile_sets(dropin, case):
    """Many antennas x beams switch the fused kernel to K-streamed B tiles (a ring of 32-antenna k-blocks, N tiles up
    to 128 columns, all time-tile accumulators open at once).  Same result as the whole-tile-set mode
    (DCBF_FLAG_DEBUG_NO_KSTREAM) to float32 rounding, and both inside the oracle budget."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m, n, xid = case
    assert _capi.fused_tiling(a, m)[2] > 1  # the shape really needs several N tiles
    x = orc.make_samples(b, a, c, t, seed=3 * a)
    dv = orc.make_delay_vals_random(c, m, a, seed=5 * m)
    dx, ddv = torch.from_numpy(x).cuda(), torch.from_numpy(dv).cuda()
    outs = []
    for flags in (0, _capi.FLAG_DEBUG_NO_KSTREAM):
        o = torch.full((b, 2, c, t // 16, 16, 2 * m), float("nan"), dtype=torch.float32, device="cuda")
        _capi.fused(dx, ddv, o, b, a, c, n, t, m, xid, TS, flags)
        _capi.fused_status()
        outs.append(o.cpu().numpy().astype(np.float64))
    ref = orc.beamform_pipeline(x, dv, n, xid, TS)
    for o in outs:
        assert not np.isnan(o).any()
        assert np.all(np.abs(o - ref) <= 2.0 ** -8 * _budget(x) + 1e-3)
    np.testing.assert_allclose(outs[0], outs[1], rtol=0, atol=float(_budget(x).max()) * 2.0 ** -8)


def test_fused_matches_three_kernel_chain_and_materialises_intermediates(dropin):
    from beamforming.beamform_op_sequence import OpSequenceTemplate

    ctx, queue = dropin
    b, a, c, t, m, n, xid = 2, 19, 12, 256, 6, 1024, 1
    x = orc.make_samples(b, a, c, t, seed=9)
    dv = orc.make_delay_vals_random(c, m, a, seed=10)
    outs = {}
    for fused in (True, False):
        op = OpSequenceTemplate(ctx, b, 2, c, n, t // 16, 16, a, m, xid, TS, t).instantiate(queue)
        op.fused = fused
        op.materialize_intermediates = True
        op.ensure_all_bound()
        op.buffer("bufin_reorder").set(queue, x)
        op.buffer("bufin_delay_vals").set(queue, dv)
        op()
        outs[fused] = op.buffer("bufout_mult").get(queue)
        np.testing.assert_array_equal(op.buffer("bufint_data").get(queue), orc.reorder(x))
        co = op.buffer("bufint_coeff").get(queue).astype(np.float64)
        ref_co = orc.steering_coeffs(dv, b, 2, c, n, a, m, xid, TS, out_dtype=np.float64)
        assert np.abs(co - ref_co).max() <= 1e-6
    np.testing.assert_allclose(outs[True], outs[False], rtol=0, atol=float(_budget(x).max()) * 2.0 ** -6)


def test_fused_linearity_and_channel_independence_full_size(dropin):
    """Size-independent properties at BASELINE configs[1] size (64 ants x 1024 chans x 256 x 16 beams).

    (1) channel independence: the fused result of a channel-sharded call (xeng_id = 1, second half) equals the
        corresponding half of the full call bit for bit;  (2) linearity in the voltages: beamform(x) for
        x = hi*16 + lo equals 16*beamform(hi) + beamform(lo) to float32 rounding;  (3) a sampled slice agrees
        with the oracle.
    """
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m = 1, 64, 1024, 256, 16
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randint(0, 256, (b, a, c, t, 2, 2), dtype=torch.uint8, device="cuda", generator=g)
    dv = torch.from_numpy(orc.make_delay_vals_random(c, m, a, seed=3)).cuda()
    full = torch.empty((b, 2, c, t // 16, 16, 2 * m), dtype=torch.float32, device="cuda")
    _capi.fused(x, dv, full, b, a, c, c, t, m, 0, TS)
    half = torch.empty((b, 2, c // 2, t // 16, 16, 2 * m), dtype=torch.float32, device="cuda")
    _capi.fused(x[:, :, c // 2:].contiguous(), dv[c // 2:].contiguous(), half, b, a, c // 2, c, t, m, 1, TS)
    _capi.fused_status()
    assert torch.equal(full[:, :, c // 2:], half)
    hi, lo = x >> 4, x & 15
    o_hi, o_lo = torch.empty_like(full), torch.empty_like(full)
    _capi.fused(hi, dv, o_hi, b, a, c, c, t, m, 0, TS)
    _capi.fused(lo, dv, o_lo, b, a, c, c, t, m, 0, TS)
    _capi.fused_status()
    assert (16 * o_hi + o_lo - full).abs().max().item() <= 0.05
    sl = slice(500, 504)
    xs, dvs = x[:, :, sl].cpu().numpy(), dv[sl].cpu().numpy()
    # channels 500..503 of a 1024-channel stream: emulate by xeng geometry C=4, xeng_id=125
    ref = orc.beamform_pipeline(xs, np.ascontiguousarray(dvs), c, 125, TS)
    assert np.all(np.abs(full[:, :, sl].cpu().numpy() - ref) <= _budget(xs))


def test_host_plan_matches_device_path(dropin):
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m, n, xid = 2, 20, 37, 64, 5, 256, 3
    x = orc.make_samples(b, a, c, t, seed=21)
    dv = orc.make_delay_vals_random(c, m, a, seed=22)
    out = np.zeros((b, 2, c, t // 16, 16, 2 * m), np.float32)
    plan = _capi.HostPlan(b, a, c, n, t, m, xid, TS, chunk_chans=8, n_slots=3)
    plan.run(x, dv, out)
    plan.close()
    dev_out = torch.empty(out.shape, dtype=torch.float32, device="cuda")
    _capi.fused(torch.from_numpy(x).cuda(), torch.from_numpy(dv).cuda(), dev_out, b, a, c, n, t, m, xid, TS)
    _capi.fused_status()
    np.testing.assert_array_equal(out, dev_out.cpu().numpy())
    assert np.all(np.abs(out - orc.beamform_pipeline(x, dv, n, xid, TS)) <= _budget(x))


def test_host_plan_resident_delay_model(dropin):
    """dcbf_host_plan_set_delay_vals: the delay model is uploaded once, runs with delay_vals = NULL use it (per-step
    H2D = voltages only) and give the bytes of the per-step-upload form; an update lands in the other copy and takes
    effect from the next run."""
    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m, n, xid = 1, 24, 29, 32, 6, 128, 2
    x = orc.make_samples(b, a, c, t, seed=71)
    dv1 = orc.make_delay_vals_random(c, m, a, seed=72)
    dv2 = orc.make_delay_vals_random(c, m, a, seed=73)
    plan = _capi.HostPlan(b, a, c, n, t, m, xid, TS, chunk_chans=7, n_slots=2)
    want1, want2, got = (np.zeros((b, 2, c, t // 16, 16, 2 * m), np.float32) for _ in range(3))
    with pytest.raises(ValueError):
        plan.run(x, None, got)  # no resident model yet
    plan.run(x, dv1, want1)
    plan.run(x, dv2, want2)
    plan.set_delay_vals(dv1)
    plan.run(x, None, got)
    np.testing.assert_array_equal(got, want1)
    plan.set_delay_vals(dv2)
    plan.run(x, None, got)
    np.testing.assert_array_equal(got, want2)
    assert not np.array_equal(want1, want2)
    plan.close()


def test_c_abi_error_codes(dropin):
    import torch

    from dpdk_dc_sand_b200 import _capi

    x = torch.zeros(64, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        _capi.reorder(x, x, 1, 1, 1, 24)  # T % 16 != 0
    with pytest.raises(ValueError):
        _capi.fused(x, x, x, 1, 0, 1, 1, 16, 1, 0, TS)  # n_ants = 0
    lib = _capi.load()
    assert lib.dcbf_reorder(None, None, 1, 1, 1, 16, None) == _capi.ERR_INVALID_ARG
    assert b"invalid" in lib.dcbf_strerror(_capi.ERR_INVALID_ARG)


def test_native_library_is_loaded():
    """The round-end harness records which in-tree .so files the test process loaded."""
    from dpdk_dc_sand_b200 import _capi

    _capi.load()
    with open("/proc/self/maps") as fh:
        assert any("libdcbf.so" in line for line in fh)
    assert os.path.basename(_capi.lib_path()) == "libdcbf.so"


def test_fused_launches_can_be_captured_in_a_cuda_graph(dropin):
    """The C ABI only enqueues (no allocation or synchronisation after the first call on a device), so a sequence of
    launches can be captured once and replayed: graph replays, interleaved with live launches on another stream,
    reproduce the eager results bit for bit (captured launches keep a channel-queue slot of their own)."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m, n = 1, 64, 300, 128, 16, 1024
    g = torch.Generator(device="cuda").manual_seed(17)
    xs = [torch.randint(0, 256, (b, a, c, t, 2, 2), dtype=torch.uint8, device="cuda", generator=g) for _ in range(3)]
    dv = torch.from_numpy(orc.make_delay_vals_random(c, m, a, seed=5)).cuda()
    outs = [torch.empty((b, 2, c, t // 16, 16, 2 * m), dtype=torch.float32, device="cuda") for _ in range(3)]
    want = []
    for x, o in zip(xs, outs):  # eager pass (also performs the one-time allocations)
        _capi.fused(x, dv, o, b, a, c, n, t, m, 0, TS)
        torch.cuda.synchronize()
        want.append(o.clone())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        st = torch.cuda.current_stream()
        for x, o in zip(xs, outs):
            _capi.fused(x, dv, o, b, a, c, n, t, m, 0, TS, 0, st)
    side = torch.cuda.Stream()
    live = torch.empty_like(outs[0])
    for rep in range(3):
        for o in outs:
            o.fill_(float("nan"))
        torch.cuda.synchronize()
        graph.replay()
        for _ in range(4):  # live launches racing the replay
            _capi.fused(xs[0], dv, live, b, a, c, n, t, m, 0, TS, 0, side)
        torch.cuda.synchronize()
        _capi.fused_status()
        for o, w in zip(outs, want):
            assert torch.equal(o, w), rep
        assert torch.equal(live, want[0])


def test_headline_size_fused_equals_three_kernel_chain_and_is_linear(dropin):
    """Size-independent properties at the size the BASELINE metric is quoted on (C3: 64 antennas x 4096 channels x
    256 samples, 64 beams): (1) the fused kernel agrees with the chain reorder -> float64 coefficients -> tcgen05
    contraction, four different kernels and two different coefficient arithmetics, to 0.05 absolute (the 2^-10 budget
    is ~8 here);  (2) the contraction is linear in the voltages;  (3) the reorder is an exact permutation (its byte
    histogram and a strided probe of elements survive)."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m = 1, 64, 4096, 256, 64
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(23)
    x = torch.randint(0, 256, (b, a, c, t, 2, 2), dtype=torch.uint8, device=dev, generator=g)
    dv = torch.zeros((c, m, a, 4), dtype=torch.float32, device=dev)
    dv[..., 0] = (torch.rand((c, m, a), device=dev, generator=g) * 32 - 16) * TS
    dv[..., 2] = (torch.rand((c, m, a), device=dev, generator=g) * 2 - 1) * math.pi
    fused = torch.empty((b, 2, c, t // 16, 16, 2 * m), dtype=torch.float32, device=dev)
    _capi.fused(x, dv, fused, b, a, c, c, t, m, 0, TS)
    re = torch.empty((b, 2, c, t // 16, 16, a, 2), dtype=torch.uint8, device=dev)
    co = torch.empty((b, 2, c, 2 * a, 2 * m), dtype=torch.float32, device=dev)
    chain = torch.empty_like(fused)
    _capi.reorder(x, re, b, a, c, t)
    _capi.coeffs(dv, co, b, 2, c, c, a, m, 0, TS)
    _capi.beamform(re, co, chain, b, c, t, a, m)
    torch.cuda.synchronize()
    _capi.fused_status()
    assert (fused - chain).abs().max().item() <= 0.05
    # reorder: (b, a, c, t, p, x) -> (b, p, c, t//16, t%16, a, x), probed on a strided subset, plus the histogram
    probe = x[:, ::7, ::513, ::5].permute(0, 4, 2, 3, 1, 5)
    got = re.reshape(b, 2, c, t, a, 2)[:, :, ::513, ::5, ::7]
    assert torch.equal(got, probe)
    assert torch.equal(torch.bincount(x.reshape(-1).int(), minlength=256), torch.bincount(re.reshape(-1).int(), minlength=256))
    del co, chain
    hi, lo = re >> 4, re & 15
    co = torch.randn((b, 2, c, 2 * a, 2 * m), dtype=torch.float32, device=dev, generator=g)
    o, o_hi, o_lo = torch.empty_like(fused), torch.empty_like(fused), torch.empty_like(fused)
    _capi.beamform(re, co, o, b, c, t, a, m)
    _capi.beamform(hi, co, o_hi, b, c, t, a, m)
    _capi.beamform(lo, co, o_lo, b, c, t, a, m)
    torch.cuda.synchronize()
    _capi.fused_status()
    assert (16 * o_hi + o_lo - o).abs().max().item() <= 0.05


# BASELINE.json configs[2..4] whole on one GPU and as the per-GPU share of 8 (xeng_id > 0): n_ants, n_chans on this GPU,
# n_samples, n_beams, n_chans of the band, xeng_id
FULL_SIZE = {
    "c3": (64, 4096, 256, 64, 4096, 0),
    "c3_share_of_8": (64, 512, 256, 64, 4096, 5),
    "c4": (80, 32768, 256, 32, 32768, 0),
    "c4_share_of_8": (80, 4096, 256, 32, 32768, 7),
    "c5": (197, 4096, 256, 256, 4096, 0),
    "c5_share_of_8": (197, 512, 256, 256, 4096, 3),
}


def _full_size_inputs(a, c, t, m, seed):
    import torch

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randint(0, 256, (1, a, c, t, 2, 2), dtype=torch.uint8, device=dev, generator=g)
    dv = torch.zeros((c, m, a, 4), dtype=torch.float32, device=dev)
    dv[..., 0] = (torch.rand((c, m, a), device=dev, generator=g) * 32 - 16) * TS
    dv[..., 2] = (torch.rand((c, m, a), device=dev, generator=g) * 2 - 1) * math.pi
    return x, dv


def _sampled_channels(c, seed, count=16):
    rng = np.random.default_rng(seed)
    return sorted({0, c - 1, *rng.choice(c, size=count - 2, replace=False).tolist()})


@pytest.mark.parametrize("name", list(FULL_SIZE))
def test_full_size_configs_against_the_oracle(dropin, name):
    """Every BASELINE configuration at its full size (whole band on one GPU: 64-bit offsets, many-N-tile and K-streamed
    modes) and as the 8-GPU share with a non-zero xeng_id: 16 channels (first, last, 14 random) of the fused result
    against the float64 oracle within the 2^-10 * sum|x| budget."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    a, c, t, m, n_total, xid = FULL_SIZE[name]
    x, dv = _full_size_inputs(a, c, t, m, seed=31)
    out = torch.empty((1, 2, c, t // 16, 16, 2 * m), dtype=torch.float32, device=x.device)
    _capi.fused(x, dv, out, 1, a, c, n_total, t, m, xid, TS)
    torch.cuda.synchronize()
    _capi.fused_status()
    worst = 0.0
    for ch in _sampled_channels(c, seed=32):
        xs = x[:, :, ch:ch + 1].cpu().numpy()
        dvs = np.ascontiguousarray(dv[ch:ch + 1].cpu().numpy())
        ref = orc.beamform_pipeline(xs, dvs, n_total, c * xid + ch, TS)  # a 1-channel engine at the absolute channel
        got = out[:, :, ch:ch + 1].cpu().numpy()
        err = np.abs(got.astype(np.float64) - ref)
        assert np.all(err <= _budget(xs)), (name, ch, float(err.max()))
        worst = max(worst, float(np.max(err / (_budget(xs) * 2.0 ** 10))))
    assert worst < 2.0 ** -10


@pytest.mark.parametrize("name", ["c3", "c3_share_of_8", "c5_share_of_8"])
def test_full_size_single_rounding_option_stays_inside_the_budget(dropin, name):
    """DCBF_FLAG_FP16_COEFF (one fp16 rounding per coefficient, one MMA pass: the option that holds 0.9 of the roofline
    at the board's power cap) with DCBF_FLAG_STREAMING at full size: sampled channels against the float64 oracle.
    The north_star budget is 2^-10 * sum|x|; a rounding of 2^-12 per coefficient must stay well inside it."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    a, c, t, m, n_total, xid = FULL_SIZE[name]
    x, dv = _full_size_inputs(a, c, t, m, seed=51)
    outs = [torch.empty((1, 2, c, t // 16, 16, 2 * m), dtype=torch.float32, device=x.device) for _ in range(2)]
    for i in range(2):  # two overlapping launches into distinct buffers, as the streaming flag requires
        _capi.fused(x, dv, outs[i], 1, a, c, n_total, t, m, xid, TS, _capi.FLAG_FP16_COEFF | _capi.FLAG_STREAMING)
    torch.cuda.synchronize()
    _capi.fused_status()
    assert torch.equal(outs[0], outs[1])
    worst = 0.0
    for ch in _sampled_channels(c, seed=52):
        xs = x[:, :, ch:ch + 1].cpu().numpy()
        dvs = np.ascontiguousarray(dv[ch:ch + 1].cpu().numpy())
        ref = orc.beamform_pipeline(xs, dvs, n_total, c * xid + ch, TS)
        err = np.abs(outs[1][:, :, ch:ch + 1].cpu().numpy().astype(np.float64) - ref)
        worst = max(worst, float(np.max(err / _budget(xs))))
    assert worst < 0.25, worst  # (measured: under a tenth of the budget)


PACKED_CASES = [
    # B, A, C, T, M, N, xeng_id, flags
    (1, 64, 300, 256, 64, 4096, 3, 0),                         # C3 geometry, more channels than SMs (cut last round)
    (1, 64, 170, 256, 16, 1024, 0, 0),                         # C2 geometry (merged hi|lo tiles, extra ring stages)
    (2, 80, 40, 256, 32, 32768, 7, 0),                         # C4 geometry, two heaps per launch
    (1, 4, 64, 256, 4, 64, 0, 0),                              # BASELINE configs[0]
    (2, 5, 3, 32, 3, 256, 1, 0),                               # odd everything (register epilogue), partial time tile
    (1, 64, 149, 16, 4, 4096, 3, 1),                           # signed input, 16-sample heaps
    (3, 33, 200, 144, 9, 4096, 2, 2),                          # single-rounding coefficients (half-size tile sets)
    (1, 256, 6, 256, 16, 1024, 0, 4),                          # 8 k-blocks, streaming launches
]


@pytest.mark.parametrize("case", PACKED_CASES, ids=lambda c: "B{}A{}C{}T{}M{}N{}x{}f{}".format(*c))
def test_packed_coefficients_give_the_same_beams(dropin, case):
    """dcbf_fused_pack_coeffs + dcbf_fused_packed (the delay model evaluated once, its tile sets loaded per heap) against
    dcbf_fused: bit-identical -- the packed bytes ARE what the kernel builds in shared memory -- and against the float64
    oracle; a second delay model packed into the same buffer replaces the first."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m, n, xid, flags = case
    dev = torch.device("cuda", 0)
    x = orc.make_samples(b, a, c, t, seed=300 + a)
    dvs = [orc.make_delay_vals_random(c, m, a, seed=400 + m + i) for i in range(2)]
    dx = torch.from_numpy(x).to(dev)
    nbytes = _capi.fused_packed_bytes(a, c, m, flags)
    assert nbytes == c * _capi.fused_tiling(a, m, flags)[0] * (1 if flags & 2 else 2) * _capi.fused_tiling(a, m, flags)[1] * 128
    packed = torch.full((nbytes + 256,), 0x5A, dtype=torch.uint8, device=dev)  # (a canary tail)
    shape = (b, 2, c, t // 16, 16, 2 * m)
    for i, dv in enumerate(dvs):
        ddv = torch.from_numpy(dv).to(dev)
        want = torch.full(shape, float("nan"), dtype=torch.float32, device=dev)
        got = torch.full(shape, float("nan"), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        _capi.fused(dx, ddv, want, b, a, c, n, t, m, xid, TS, flags)
        _capi.fused_pack_coeffs(ddv, packed, a, c, n, m, xid, TS, flags & _capi.FLAG_FP16_COEFF)
        if flags & _capi.FLAG_STREAMING:  # a streaming launch promises independence of the kernel queued before it:
            torch.cuda.synchronize()      # the first one after a pack is not (dcbf.h), so the pack is waited for here
        n0 = _capi.launch_count()
        _capi.fused_packed(dx, packed, got, b, a, c, n, t, m, xid, TS, flags)
        torch.cuda.synchronize()
        _capi.fused_status()
        assert _capi.launch_count() - n0 == 1
        assert torch.equal(got, want), (i, int((got != want).sum()))
        assert bool((packed[nbytes:] == 0x5A).all())
        if i == 0:
            ref = orc.beamform_pipeline(x, dv, n, xid, TS, signed_input=bool(flags & 1))
            assert np.all(np.abs(got.cpu().numpy().astype(np.float64) - ref) <= _budget(x, bool(flags & 1)))


@pytest.mark.parametrize("case", [(1, 64, 300, 256, 64, 0), (2, 64, 170, 256, 16, 0), (1, 80, 40, 256, 32, 0), (2, 5, 7, 48, 3, 0),
                                  (1, 33, 200, 144, 8, 2), (1, 1, 20, 64, 16, 1)],
                         ids=lambda c: "B{}A{}C{}T{}M{}f{}".format(*c))
def test_packed_coefficients_int8_output(dropin, case):
    """dcbf_fused_pack_coeffs_q8 + dcbf_fused_packed_q8 against dcbf_fused_q8: the same int8 beams and the same
    saturation count (the gains ride on the packed coefficients), and the oracle's requantisation within one step."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    b, a, c, t, m, flags = case
    dev = torch.device("cuda", 0)
    n, xid = 2 * c, 1
    x = orc.make_samples(b, a, c, t, seed=500 + a)
    dv = orc.make_delay_vals_random(c, m, a, seed=600 + m)
    gains = np.linspace(0.004, 0.03, m).astype(np.float32)
    dx, ddv, dg = (torch.from_numpy(v).to(dev) for v in (x, dv, gains))
    want = torch.zeros((b, 2, c, t // 16, 16, 2 * m), dtype=torch.int8, device=dev)
    got = torch.full_like(want, 77)
    sat_w, sat_g = (torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(2))
    packed = torch.empty(_capi.fused_packed_bytes(a, c, m, flags), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    _capi.fused_q8(dx, ddv, dg, want, b, a, c, n, t, m, xid, TS, flags, saturated=sat_w)
    _capi.fused_pack_coeffs_q8(ddv, dg, packed, a, c, n, m, xid, TS, flags & _capi.FLAG_FP16_COEFF)
    _capi.fused_packed_q8(dx, packed, dg, got, b, a, c, n, t, m, xid, TS, flags, saturated=sat_g)
    torch.cuda.synchronize()
    _capi.fused_status()
    assert torch.equal(got, want)
    assert int(sat_g.item()) == int(sat_w.item())
    ref, _ = orc.requantise(orc.beamform_pipeline(x, dv, n, xid, TS, signed_input=bool(flags & 1)), gains)
    assert np.abs(got.cpu().numpy().astype(np.int32) - ref.astype(np.int32)).max() <= 1


def test_packed_coefficients_int8_operator(dropin):
    """QuantisedOpSequence.pack_coefficients(): same int8 beams and saturation count as the per-call path."""
    from beamforming.beamform_op_sequence import QuantisedOpSequenceTemplate

    ctx, queue = dropin
    b, a, c, t, m, n, xid = 1, 64, 20, 256, 64, 4096, 1
    op = QuantisedOpSequenceTemplate(ctx, b, 2, c, n, t // 16, 16, a, m, xid, TS, t).instantiate(queue)
    op.ensure_all_bound()
    op.buffer("bufin_reorder").set(queue, orc.make_samples(b, a, c, t, seed=9))
    op.buffer("bufin_delay_vals").set(queue, orc.make_delay_vals_random(c, m, a, seed=10))
    op.buffer("bufin_gains").set(queue, np.full(m, 0.03, np.float32))
    op()
    want, sat = op.buffer("bufout_q8").get(queue).copy(), op.saturated
    assert op.pack_coefficients()
    op.buffer("bufout_q8").zero(queue)
    op()
    assert np.array_equal(op.buffer("bufout_q8").get(queue), want)
    assert op.saturated == sat and sat > 0


def test_packed_coefficients_through_the_operator_and_unsupported_shapes(dropin):
    """OpSequence.pack_coefficients(): the calls after it load the packed tile sets (one launch each, same beams as
    before); a new delay model needs a new pack; shapes without a whole tile set stay on the per-call path."""
    import torch

    from beamforming.beamform_op_sequence import OpSequenceTemplate
    from dpdk_dc_sand_b200 import _capi

    ctx, queue = dropin
    b, a, c, t, m, n, xid = 1, 64, 40, 256, 64, 4096, 2
    op = OpSequenceTemplate(ctx, b, 2, c, n, t // 16, 16, a, m, xid, TS, t).instantiate(queue)
    op.ensure_all_bound()
    x = orc.make_samples(b, a, c, t, seed=5)
    dv = orc.make_delay_vals_random(c, m, a, seed=6)
    op.buffer("bufin_reorder").set(queue, x)
    op.buffer("bufin_delay_vals").set(queue, dv)
    op()
    want = op.buffer("bufout_mult").get(queue).copy()
    assert op.pack_coefficients()
    op.buffer("bufout_mult").zero(queue)
    n0 = _capi.launch_count()
    op()
    got = op.buffer("bufout_mult").get(queue)
    assert _capi.launch_count() - n0 == 1
    assert np.array_equal(got, want)
    dv2 = orc.make_delay_vals_random(c, m, a, seed=7)
    op.buffer("bufin_delay_vals").set(queue, dv2)
    op()  # still the packed (old) model: the delay model is only read when it is packed
    assert np.array_equal(op.buffer("bufout_mult").get(queue), want)
    op.pack_coefficients()
    op()
    ref = orc.beamform_pipeline(x, dv2, n, xid, TS)
    assert np.all(np.abs(op.buffer("bufout_mult").get(queue).astype(np.float64) - ref) <= _budget(x))
    op.release_coefficients()
    op()
    assert np.all(np.abs(op.buffer("bufout_mult").get(queue).astype(np.float64) - ref) <= _budget(x))
    _capi.fused_status()
    # many antennas x beams: K-streamed, no whole tile set
    assert _capi.fused_packed_bytes(197, 8, 256) == 0
    dev = torch.device("cuda", 0)
    dummy = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
    with pytest.raises(Exception):
        _capi.fused_pack_coeffs(torch.zeros((8, 256, 197, 4), device=dev), dummy, 197, 8, 8, 256, 0, TS)
    big = OpSequenceTemplate(ctx, 1, 2, 2, 64, 8, 16, 197, 256, 0, TS, 128).instantiate(queue)
    big.ensure_all_bound()
    assert big.pack_coefficients() is False


@pytest.mark.parametrize("shape", [(64, 1024, 256, 16), (64, 512, 256, 64), (80, 600, 256, 32), (197, 64, 256, 256)],
                         ids=["c2", "c3_share", "c4_like", "c5_like"])
def test_repeated_launches_are_bit_identical(dropin, shape):
    """200 launches queued back to back over rotating input / output sets (programmatic dependent launch lets each start
    while its predecessor drains) must all reproduce the first result of their input set bit for bit: a tile overwritten
    before its last reader is done would show here and not in a test that launches once (tools/stress_determinism.py is
    the long version)."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    a, c, t, m = shape
    dev = torch.device("cuda", 0)
    sets, reps = 3, 200
    inputs = [_full_size_inputs(a, c, t, m, seed=61 + j) for j in range(sets)]
    out_shape = (1, 2, c, t // 16, 16, 2 * m)
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()  # the inputs were generated on the default stream, the launches go to `stream`
    refs = []
    for x, dv in inputs:
        o = torch.empty(out_shape, dtype=torch.float32, device=dev)
        _capi.fused(x, dv, o, 1, a, c, c, t, m, 0, TS, 0, stream)
        stream.synchronize()
        refs.append(o)
    ring = [torch.full(out_shape, float("nan"), dtype=torch.float32, device=dev) for _ in range(8)]
    torch.cuda.synchronize()
    bad, detail = 0, ""
    for start in range(0, reps, len(ring)):
        for i in range(len(ring)):
            x, dv = inputs[(start + i) % sets]
            _capi.fused(x, dv, ring[i], 1, a, c, c, t, m, 0, TS, 0, stream)
        stream.synchronize()
        for i in range(len(ring)):
            ref = refs[(start + i) % sets]
            if not torch.equal(ring[i], ref):
                bad += 1
                if not detail:
                    diff = (ring[i] != ref) | torch.isnan(ring[i])
                    idx = diff.nonzero()
                    detail = (f"launch {start + i}: {int(diff.sum())} values differ, nan {int(torch.isnan(ring[i]).sum())}, "
                              f"max |d| {float((ring[i] - ref).abs().nan_to_num().max()):.3g}, first {idx[0].tolist()}, last {idx[-1].tolist()}, "
                              f"channels {sorted(set(idx[:, 2].tolist()))[:12]}, pols {sorted(set(idx[:, 1].tolist()))}, "
                              f"columns {sorted(set(idx[:, 5].tolist()))[:8]}...")
    _capi.fused_status()
    assert bad == 0, f"{bad} of {reps} launches differ; {detail}"


def test_full_size_q8_against_the_oracle(dropin):
    """int8 requantised output at the size the metric is quoted on (C3): sampled channels against the oracle's
    requantisation of its float64 beams (at most one quantisation step away, practically always equal)."""
    import torch

    from dpdk_dc_sand_b200 import _capi

    a, c, t, m, n_total, xid = FULL_SIZE["c3"]
    x, dv = _full_size_inputs(a, c, t, m, seed=41)
    gains = torch.full((m,), 0.02, dtype=torch.float32, device=x.device)  # sum|x| ~ 1e4: a few per cent of the values clip
    out = torch.empty((1, 2, c, t // 16, 16, 2 * m), dtype=torch.int8, device=x.device)
    sat = torch.zeros(1, dtype=torch.int64, device=x.device)
    _capi.fused_q8(x, dv, gains, out, 1, a, c, n_total, t, m, xid, TS, saturated=sat)
    torch.cuda.synchronize()
    _capi.fused_status()
    differing = total = 0
    for ch in _sampled_channels(c, seed=42):
        xs = x[:, :, ch:ch + 1].cpu().numpy()
        dvs = np.ascontiguousarray(dv[ch:ch + 1].cpu().numpy())
        want, _ = orc.requantise(orc.beamform_pipeline(xs, dvs, n_total, c * xid + ch, TS), gains.cpu().numpy())
        diff = np.abs(out[:, :, ch:ch + 1].cpu().numpy().astype(np.int32) - want.astype(np.int32))
        assert diff.max() <= 1, ch
        differing += int(np.count_nonzero(diff))
        total += diff.size
    assert differing <= 1e-3 * total
    assert 0 < int(sat.item()) < out.numel()
