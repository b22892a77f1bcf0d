"""CPU oracle for the tied-array beamforming hot path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference's CPU implementations of
the three stages of the path (reorder -> steering coefficients -> contraction
over antennas).  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
Nothing under ``dpdk_dc_sand_b200/`` imports it and the product path has no CPU
fallback.

Parity is PINNED: ``tests/golden/make_golden.py`` runs the reference's own
functions (imported from ``/root/reference/beamformer``) on small seeded inputs
and commits their outputs under ``tests/golden/``; ``tests/test_oracle.py``
checks every function here against those fixtures.

Reference lines restated (paths relative to the reference root):

* reorder            ``beamformer/beamforming/reorder.py:40-42``
* coefficients       ``beamformer/unit_test/coeff_generator_cpu.py:120-186``
                     (same formula as ``beamformer/beamforming/coeff_generator.py:49-103``
                     but with the consistent ``[c][beam][ant]`` read index)
* contraction        ``beamformer/beamforming/complex_mult_kernel.py:89-100``
                     (the K3 definition over all ``2*n_beams`` columns)
* beam-0 shortcut    ``beamformer/unit_test/complex_mult_cpu.py:76-102``
                     (the reference CPU checker re-uses beam 0's coefficients for
                     every beam; restated separately as ``complex_mult_beam0``)

Arithmetic type of the coefficient formula.  The reference was written for
numpy 1.x, where ``np.float32_scalar * python_int`` promotes to float64, and its
numba GPU kernel also promotes ``float32 * int64`` to float64; both store the
result as float32.  Under numpy >= 2 the same CPU source silently evaluates in
float32 (NEP 50).  The oracle restates the *float64* evaluation (the north-star
tolerance is "1e-6 of a float64 evaluation") and casts explicitly so that the
numpy version does not matter.  ``steering_coeffs(..., f32_arith=True)`` gives
the numpy>=2 behaviour for completeness; both are pinned by golden vectors.
"""
from __future__ import annotations

import math

import numpy as np

N_POLS = 2
COMPLEXITY = 2
SAMPLES_PER_BLOCK = 16  # reference: prebeamform_reorder.py:59 (128 // 8)


# --------------------------------------------------------------------------------------
# Stage 1: pre-beamform reorder
# --------------------------------------------------------------------------------------
def reorder(samples: np.ndarray, samples_per_block: int = SAMPLES_PER_BLOCK) -> np.ndarray:
    """(B, A, C, T, P, 2) u8 -> (B, P, C, T//S, S, A, 2) u8.  reorder.py:40-42."""
    b, a, c, t, p, x = samples.shape
    if t % samples_per_block:
        raise ValueError("n_samples_per_channel must be a multiple of samples_per_block")
    v = samples.reshape(b, a, c, t // samples_per_block, samples_per_block, p, x)
    return np.ascontiguousarray(v.transpose(0, 5, 2, 3, 4, 1, 6))


# --------------------------------------------------------------------------------------
# Stage 2: steering coefficients
# --------------------------------------------------------------------------------------
def rotation(
    delay_vals: np.ndarray,
    n_channels_per_stream: int,
    n_channels: int,
    xeng_id: int,
    sample_period: float,
    f32_arith: bool = False,
    dt: float = None,
) -> np.ndarray:
    """Rotation angle (radians) per (c, beam, ant).  coeff_generator_cpu.py:125-165.

    Operation order is kept identical to the reference expression so that the
    float64 result is bit-identical to the reference loop.

    ``dt`` (seconds since the delay model's reference time; None = the reference Python path, which ignores
    the rate fields) selects the time-varying form of the native precursor
    (beamformer_coefficient_generator/BeamformerKernels.cu:25-35): delay -> delay + delay_rate*dt,
    phase -> phase + phase_rate*dt, evaluated in float64.  (The precursor's ``fDelayN`` line adds the delay
    RATE where the delay is evidently meant, :31; the intended formula is restated here.)
    """
    ft = np.float32 if f32_arith else np.float64
    delay_s = delay_vals[..., 0].astype(ft)  # (C, M, A)
    phase_rad = delay_vals[..., 2].astype(ft)
    # delay_vals[..., 1] (delay rate) and [..., 3] (phase rate) are ignored by the reference Python path.
    if dt is not None:
        if f32_arith:
            raise ValueError("time-varying steering is defined in float64 only")
        delay_s = delay_s + delay_vals[..., 1].astype(np.float64) * float(dt)
        phase_rad = phase_rad + delay_vals[..., 3].astype(np.float64) * float(dt)
    ichannel = (np.arange(n_channels_per_stream) + n_channels_per_stream * xeng_id).astype(ft)
    ichannel = ichannel[:, None, None]
    if f32_arith:
        # numpy>=2 semantics: python scalars are weak, every op rounds to float32.
        neg_pi = np.float32(-math.pi)
        denom = np.float32(n_channels * sample_period)
        half = np.float32(n_channels / 2)
    else:
        neg_pi = -math.pi
        denom = n_channels * sample_period
        half = n_channels / 2
    initial_phase = delay_s * ichannel * neg_pi / denom + phase_rad
    band_centre = delay_s * half * neg_pi / denom
    return initial_phase - band_centre


def steering_coeffs(
    delay_vals: np.ndarray,
    n_batches: int,
    n_pols: int,
    n_channels_per_stream: int,
    n_channels: int,
    n_ants: int,
    n_beams: int,
    xeng_id: int,
    sample_period: float,
    f32_arith: bool = False,
    out_dtype=np.float32,
    batch_dt=None,
    _dt=None,
    weights=None,
) -> np.ndarray:
    """(C, M, A, 4) f32 -> (B, P, C, 2A, 2M).  coeff_generator_cpu.py:120-186.

    Block for (ant a, beam m): rows 2a,2a+1 / cols 2m,2m+1 = [[cos, sin], [-sin, cos]].
    ``batch_dt`` (n_batches seconds, see ``rotation``) makes the coefficients of each batch its own.
    """
    if delay_vals.shape != (n_channels_per_stream, n_beams, n_ants, 4):
        raise ValueError(f"delay_vals shape {delay_vals.shape} mismatch")
    if batch_dt is not None:
        if len(batch_dt) != n_batches:
            raise ValueError("batch_dt needs one entry per batch")
        per_batch = [
            steering_coeffs(delay_vals, 1, n_pols, n_channels_per_stream, n_channels, n_ants, n_beams, xeng_id,
                            sample_period, out_dtype=out_dtype, batch_dt=None, _dt=float(t), weights=weights)
            for t in batch_dt
        ]
        return np.concatenate(per_batch, axis=0)
    rot = rotation(delay_vals, n_channels_per_stream, n_channels, xeng_id, sample_period, f32_arith, dt=_dt)
    rot = rot.astype(np.float64)  # math.cos/math.sin take a C double
    cos = np.cos(rot).transpose(0, 2, 1)  # (C, A, M)
    sin = np.sin(rot).transpose(0, 2, 1)
    if weights is not None:  # real weight per (beam, antenna): the ?beam-weights request (ngkcs/ngkcs/corr3_servlet.py:140)
        w = np.asarray(weights, np.float64).T[None]  # (1, A, M)
        cos, sin = cos * w, sin * w
    blk = np.empty((n_channels_per_stream, n_ants, 2, n_beams, 2), dtype=np.float64)
    blk[:, :, 0, :, 0] = cos
    blk[:, :, 0, :, 1] = sin
    blk[:, :, 1, :, 0] = -sin
    blk[:, :, 1, :, 1] = cos
    blk = blk.reshape(n_channels_per_stream, 2 * n_ants, 2 * n_beams).astype(out_dtype)
    return np.ascontiguousarray(np.broadcast_to(blk, (n_batches, n_pols) + blk.shape))


def steering_coeffs_loop(
    delay_vals, n_batches, n_pols, n_channels_per_stream, n_channels, n_ants, n_beams, xeng_id, sample_period
) -> np.ndarray:
    """Scalar-loop restatement (libm cos/sin, float64) for small cases; pins the vectorised form."""
    out = np.empty((n_batches, n_pols, n_channels_per_stream, 2 * n_ants, 2 * n_beams), np.float32)
    for c in range(n_channels_per_stream):
        ch = c + n_channels_per_stream * xeng_id
        for m in range(n_beams):
            for a in range(n_ants):
                d = float(delay_vals[c, m, a, 0])
                ph = float(delay_vals[c, m, a, 2])
                initial = d * ch * (-math.pi) / (n_channels * sample_period) + ph
                centre = d * (n_channels / 2) * (-math.pi) / (n_channels * sample_period)
                r = initial - centre
                re, im = math.cos(r), math.sin(r)
                out[:, :, c, 2 * a, 2 * m] = re
                out[:, :, c, 2 * a, 2 * m + 1] = im
                out[:, :, c, 2 * a + 1, 2 * m] = -im
                out[:, :, c, 2 * a + 1, 2 * m + 1] = re
    return out


# --------------------------------------------------------------------------------------
# Stage 3: contraction over antennas
# --------------------------------------------------------------------------------------
def _as_real(reordered: np.ndarray, signed_input: bool, dtype) -> np.ndarray:
    b, p, c, k, s, a, x = reordered.shape
    d = reordered.reshape(b, p, c, k, s, a * x)
    if signed_input:
        d = d.view(np.int8)
    return d.astype(dtype)


def beamform(
    reordered: np.ndarray, coeffs: np.ndarray, signed_input: bool = False, acc_dtype=np.float64
) -> np.ndarray:
    """K3 definition: out[b,p,c,k,s,col] = sum_j f(data[b,p,c,k,s,j]) * coeff[b,p,c,j,col].

    complex_mult_kernel.py:89-100 (data viewed as [..., 2A], :127-134).  Bytes are
    unsigned in the reference API (matrix_multiply.py:146); ``signed_input`` is the
    int8 interpretation the F-engine actually produces.  Evaluated in ``acc_dtype``
    (float64 = ground truth; float32 = same working precision as the reference).
    """
    d = _as_real(reordered, signed_input, acc_dtype)
    out = np.einsum("bpcksj,bpcjn->bpcksn", d, coeffs.astype(acc_dtype), optimize=True)
    return out


def beamform_abs_bound(reordered: np.ndarray, signed_input: bool = False) -> np.ndarray:
    """sum_a |x_a| per output sample (B,P,C,K,S): the north-star error budget is 2^-10 times this."""
    b, p, c, k, s, a, x = reordered.shape
    d = reordered.view(np.int8) if signed_input else reordered
    d = d.astype(np.float64)
    return np.sqrt(d[..., 0] ** 2 + d[..., 1] ** 2).sum(axis=-1)


def complex_mult_beam0(reordered: np.ndarray, coeffs: np.ndarray) -> np.ndarray:
    """Restatement of the reference CPU checker INCLUDING its beam-0 shortcut.

    complex_mult_cpu.py:76-102: for every beam the checker builds the 2A x 2 matrix
    from ``coeffs[b,p,c,2a,0]`` / ``coeffs[b,p,c,2a,1]`` (beam 0), float32 ``np.dot``.
    Only equal to ``beamform`` when coefficients are beam-uniform.
    """
    b, p, c, k, s, a, x = reordered.shape
    n_beams = coeffs.shape[4] // 2
    d = _as_real(reordered, False, np.float32)
    re0 = coeffs[:, :, :, 0::2, 0]  # (B,P,C,A)
    im0 = coeffs[:, :, :, 0::2, 1]
    m = np.empty((b, p, c, 2 * a, 2), np.float32)
    m[:, :, :, 0::2, 0] = re0
    m[:, :, :, 0::2, 1] = im0
    m[:, :, :, 1::2, 0] = -im0
    m[:, :, :, 1::2, 1] = re0
    prod = np.einsum("bpcksj,bpcjn->bpcksn", d, m)
    return np.ascontiguousarray(np.tile(prod, (1, 1, 1, 1, 1, n_beams))).astype(np.float32)


# --------------------------------------------------------------------------------------
# Whole path (what OpSequence.__call__ computes; beamform_op_sequence.py:141-154)
# --------------------------------------------------------------------------------------
def beamform_pipeline(
    samples: np.ndarray,
    delay_vals: np.ndarray,
    n_channels: int,
    xeng_id: int,
    sample_period: float,
    signed_input: bool = False,
    acc_dtype=np.float64,
    batch_dt=None,
    weights=None,
    sample_dt=None,
    tile=None,
) -> np.ndarray:
    """(B,A,C,T,P,2) u8 + (C,M,A,4) f32 -> (B,P,C,T//16,16,2M) in ``acc_dtype``.

    ``sample_dt`` (with ``batch_dt`` = time of each heap's first sample): the steering follows the delay model INSIDE
    the heap, sample t of heap b being steered at ``batch_dt[b] + t * sample_dt`` -- what the native precursor's
    per-timestamp coefficients do (beamformer_coefficient_generator/BeamformerKernels.cu:153-167, dt = t *
    SAMPLING_PERIOD * FFT_SIZE).  ``tile`` = n restates dcbf_fused_ex's approximation of it instead: one coefficient
    set per n consecutive samples, evaluated at their centre ``batch_dt[b] + (t0 + (n_t - 1) / 2) * sample_dt``."""
    b, a, c, t, p, x = samples.shape
    m = delay_vals.shape[1]
    re = reorder(samples)
    d = _as_real(re, signed_input, acc_dtype)
    if batch_dt is not None and sample_dt:
        dt_flat = d.reshape(b, p, c, t, 2 * a)
        out = np.empty((b, p, c, t, 2 * m), acc_dtype)
        step = int(tile) if tile else 1
        for ib in range(b):
            for t0 in range(0, t, step):
                n_t = min(step, t - t0)
                when = float(batch_dt[ib]) + (t0 + 0.5 * (n_t - 1)) * float(sample_dt)
                co = steering_coeffs(delay_vals, 1, 1, c, n_channels, a, m, xeng_id, sample_period,
                                     out_dtype=np.float64, _dt=when, weights=weights)[0, 0].astype(acc_dtype)
                out[ib, :, :, t0:t0 + n_t] = np.einsum("pctj,cjn->pctn", dt_flat[ib, :, :, t0:t0 + n_t], co)
        return out.reshape(b, p, c, t // SAMPLES_PER_BLOCK, SAMPLES_PER_BLOCK, 2 * m)
    if batch_dt is not None:
        co = steering_coeffs(delay_vals, b, 1, c, n_channels, a, m, xeng_id, sample_period, out_dtype=np.float64,
                             batch_dt=batch_dt, weights=weights)
        return np.einsum("bpcksj,bcjn->bpcksn", d, co[:, 0].astype(acc_dtype), optimize=True)
    co = steering_coeffs(delay_vals, 1, 1, c, n_channels, a, m, xeng_id, sample_period, out_dtype=np.float64,
                         weights=weights)
    return np.einsum("bpcksj,cjn->bpcksn", d, co[0, 0].astype(acc_dtype), optimize=True)


def beamform_pipeline_fast(samples, delay_vals, n_channels, xeng_id, sample_period, threads_hint=None):
    """Vectorised float32 port used as the timed CPU baseline (bench.py cpu_baseline).

    Same maths as ``beamform_pipeline`` with float32 BLAS matmul (all host cores numpy's
    BLAS will use) instead of float64 einsum.  Not used for parity.
    """
    b, a, c, t, p, x = samples.shape
    m = delay_vals.shape[1]
    re = reorder(samples)  # (B,P,C,K,S,A,2)
    co = steering_coeffs(delay_vals, 1, 1, c, n_channels, a, m, xeng_id, sample_period)[0, 0]  # (C,2A,2M) f32
    d = re.reshape(b, p, c, t, 2 * a).astype(np.float32)
    out = np.matmul(d, co[None, None])  # (B,P,C,T,2M)
    return out.reshape(b, p, c, t // SAMPLES_PER_BLOCK, SAMPLES_PER_BLOCK, 2 * m)


def requantise(beams: np.ndarray, gains: np.ndarray):
    """Beam post-stage of dcbf_fused_q8 (our definition; the reference stops at float32 beams):
    int8(clip(rint(beam * gain[m]), -127, 127)) with round-half-even, plus the number of clipped values.
    ``beams`` (..., 2M) float, ``gains`` (M,)."""
    g = np.repeat(np.asarray(gains, np.float64), 2)
    v = np.asarray(beams, np.float64) * g
    clipped = int(np.count_nonzero(np.abs(v) > 127.0))
    return np.clip(np.rint(v), -127, 127).astype(np.int8), clipped


# --------------------------------------------------------------------------------------
# Synthetic inputs shared by tests / bench
# --------------------------------------------------------------------------------------
SAMPLE_PERIOD = 1 / 1712e6  # beamform_coeff_test.py:79


def make_samples(n_batches, n_ants, n_chans, n_samples, seed=2021) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n_batches, n_ants, n_chans, n_samples, N_POLS, COMPLEXITY), dtype=np.uint8)


def make_delay_vals_uniform(n_chans, n_beams, n_ants, samples_delay=5, phase=math.pi / 2) -> np.ndarray:
    """The reference tests' inputs: beamform_coeff_test.py:86-90."""
    dv = np.zeros((n_chans, n_beams, n_ants, 4), np.float32)
    dv[..., 0] = np.single(samples_delay * SAMPLE_PERIOD)
    dv[..., 2] = np.single(phase)
    return dv


def make_delay_vals_random(n_chans, n_beams, n_ants, seed=2022, max_delay_samples=16.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    dv = np.zeros((n_chans, n_beams, n_ants, 4), np.float32)
    dv[..., 0] = (rng.uniform(-max_delay_samples, max_delay_samples, dv.shape[:3]) * SAMPLE_PERIOD).astype(np.float32)
    dv[..., 2] = rng.uniform(-math.pi, math.pi, dv.shape[:3]).astype(np.float32)
    # rates are populated with junk on purpose: the path must ignore them.
    dv[..., 1] = rng.standard_normal(dv.shape[:3]).astype(np.float32)
    dv[..., 3] = rng.standard_normal(dv.shape[:3]).astype(np.float32)
    return dv
