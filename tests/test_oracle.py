"""Pin the oracle (oracle/beamform_oracle.py) to the reference's own CPU outputs (tests/golden)."""
import os

import numpy as np
import pytest

from oracle import beamform_oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("tag", ["small", "t256"])
def test_reorder_matches_reference(golden_dir, tag):
    g = _load(golden_dir, f"golden_reorder_{tag}.npz")
    np.testing.assert_array_equal(orc.reorder(g["samples"]), g["reordered"])


@pytest.mark.parametrize("tag", ["uniform", "random"])
def test_coeffs_match_reference_float64(golden_dir, tag):
    g = _load(golden_dir, f"golden_coeffs_{tag}.npz")
    b, p, c, n, a, m, xid = (int(v) for v in g["params"])
    got = orc.steering_coeffs(g["delay_vals"], b, p, c, n, a, m, xid, orc.SAMPLE_PERIOD)
    # vectorised numpy cos/sin vs libm: identical after rounding to float32 on these vectors
    np.testing.assert_array_equal(got, g["coeffs_f64"])
    loop = orc.steering_coeffs_loop(g["delay_vals"], b, p, c, n, a, m, xid, orc.SAMPLE_PERIOD)
    np.testing.assert_array_equal(loop, g["coeffs_f64"])


@pytest.mark.parametrize("tag", ["uniform", "random"])
def test_coeffs_match_reference_numpy2_float32(golden_dir, tag):
    g = _load(golden_dir, f"golden_coeffs_{tag}.npz")
    b, p, c, n, a, m, xid = (int(v) for v in g["params"])
    got = orc.steering_coeffs(g["delay_vals"], b, p, c, n, a, m, xid, orc.SAMPLE_PERIOD, f32_arith=True)
    np.testing.assert_array_equal(got, g["coeffs_f32"])
    # and the two evaluations differ by float32 phase rounding only (documents why f64 is the oracle)
    assert np.abs(g["coeffs_f32"].astype(np.float64) - g["coeffs_f64"]).max() < 2e-5


def test_coeff_block_structure():
    dv = orc.make_delay_vals_random(3, 4, 5)
    co = orc.steering_coeffs(dv, 1, 2, 3, 64, 5, 4, 1, orc.SAMPLE_PERIOD)
    re, im = co[..., 0::2, 0::2], co[..., 0::2, 1::2]
    np.testing.assert_array_equal(co[..., 1::2, 0::2], -im)
    np.testing.assert_array_equal(co[..., 1::2, 1::2], re)
    np.testing.assert_allclose(re.astype(np.float64) ** 2 + im.astype(np.float64) ** 2, 1.0, atol=3e-7)
    np.testing.assert_array_equal(co[0, 0], co[0, 1])  # replicated over batch / pol


def test_pipeline_matches_reference_checker_uniform(golden_dir):
    g = _load(golden_dir, "golden_pipeline_uniform.npz")
    b, a, c, t, m, n, xid = (int(v) for v in g["params"])
    np.testing.assert_array_equal(orc.reorder(g["samples"]), g["reordered"])
    co = orc.steering_coeffs(g["delay_vals"], b, 2, c, n, a, m, xid, orc.SAMPLE_PERIOD)
    np.testing.assert_array_equal(co, g["coeffs"])
    # the reference tolerance for this comparison: beamform_op_sequence_test.py:198
    out = orc.beamform(g["reordered"], co)
    np.testing.assert_allclose(out, g["beams"], rtol=1e-4, atol=1e-4)
    out2 = orc.beamform_pipeline(g["samples"], g["delay_vals"], n, xid, orc.SAMPLE_PERIOD)
    np.testing.assert_allclose(out2, g["beams"], rtol=1e-4, atol=1e-4)
    fast = orc.beamform_pipeline_fast(g["samples"], g["delay_vals"], n, xid, orc.SAMPLE_PERIOD)
    np.testing.assert_allclose(fast, g["beams"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(orc.complex_mult_beam0(g["reordered"], co), g["beams"], rtol=1e-5, atol=1e-3)


def test_reference_checker_beam0_quirk_is_restated(golden_dir):
    """With non-uniform delays the reference checker differs from the K3 definition except on beam 0."""
    g = _load(golden_dir, "golden_pipeline_random.npz")
    co, re = g["coeffs"], g["reordered"]
    quirk = orc.complex_mult_beam0(re, co)
    np.testing.assert_allclose(quirk, g["beams_beam0_checker"], rtol=1e-5, atol=1e-3)
    k3 = orc.beamform(re, co)
    np.testing.assert_allclose(k3[..., 0:2], g["beams_beam0_checker"][..., 0:2], rtol=1e-5, atol=1e-3)
    assert np.abs(k3[..., 2:] - g["beams_beam0_checker"][..., 2:]).max() > 1.0


def test_beamform_linearity_and_signed():
    x = orc.make_samples(1, 6, 2, 32, seed=7)
    dv = orc.make_delay_vals_random(2, 3, 6, seed=8)
    u = orc.beamform_pipeline(x, dv, 128, 0, orc.SAMPLE_PERIOD)
    s = orc.beamform_pipeline(x, dv, 128, 0, orc.SAMPLE_PERIOD, signed_input=True)
    # unsigned = signed + 256 * [byte >= 128]
    hi = ((x >= 128).astype(np.uint8))
    corr = orc.beamform_pipeline(hi, dv, 128, 0, orc.SAMPLE_PERIOD)
    np.testing.assert_allclose(u, s + 256.0 * corr, rtol=0, atol=1e-9)
    bound = orc.beamform_abs_bound(orc.reorder(x))
    assert bound.shape == u.shape[:-1]
    assert np.all(np.abs(u) <= bound[..., None] * (1 + 1e-12) + 1e-9)


def test_reorder_rejects_bad_t():
    with pytest.raises(ValueError):
        orc.reorder(np.zeros((1, 2, 1, 24, 2, 2), np.uint8))
