"""Generate the golden vectors that pin ``oracle/`` to the reference's own CPU code.

Run in the authoring container only (``/root/reference`` is not on the GPU box):

    python tests/golden/make_golden.py

It imports the *unmodified* reference modules

    beamformer/beamforming/reorder.py            (numba reorder)
    beamformer/unit_test/coeff_generator_cpu.py  (pure-python cpu_coeffs)
    beamformer/unit_test/complex_mult_cpu.py     (numba complex_mult)

runs them on small seeded inputs and stores inputs + outputs in ``golden_*.npz``.
The only shim is ``np.math = math`` (``np.math`` was removed in numpy 2;
coeff_generator_cpu.py:148).  For the coefficient generator two evaluations are
stored: ``f64`` (delay_vals passed as float64, which reproduces the numpy-1.x /
numba float64 arithmetic the reference was written against) and ``f32`` (delay_vals
float32 under numpy >= 2, where the same source evaluates in float32).
"""
import math
import os
import sys

import numpy as np

np.math = math  # numpy>=2 shim, see module docstring
sys.path.insert(0, "/root/reference/beamformer")

from beamforming import reorder as ref_reorder  # noqa: E402
from unit_test import complex_mult_cpu as ref_mult  # noqa: E402
from unit_test.coeff_generator_cpu import CoeffGenerator as RefCoeffGenerator  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TS = 1 / 1712e6


def ref_coeffs(dv, b, p, c, n, a, m, xeng_id):
    return RefCoeffGenerator(dv, b, p, c, n, 16, 16, a, m, xeng_id, TS).cpu_coeffs()


def main():
    rng = np.random.default_rng(2021)

    # ---- reorder: odd antenna count, 2 batches, T=32 and T=256 -------------------------
    for tag, (b, a, c, t) in {"small": (2, 5, 3, 32), "t256": (1, 4, 2, 256)}.items():
        x = rng.integers(0, 256, (b, a, c, t, 2, 2), dtype=np.uint8)
        out_shape = (b, 2, c, t // 16, 16, a, 2)
        y = ref_reorder.reorder(x, x.shape, out_shape)
        np.savez_compressed(os.path.join(HERE, f"golden_reorder_{tag}.npz"), samples=x, reordered=y)

    # ---- coefficients: (1) the reference test's uniform inputs, (2) random per (c,m,a) ----
    cases = {}
    c, m, a, n, xid = 4, 2, 4, 1024, 0  # beamform_coeff_test.py geometry: C = N // A // 4 (here shrunk)
    dv = np.zeros((c, m, a, 4), np.float32)
    dv[..., 0] = np.single(5 * TS)
    dv[..., 2] = np.single(np.pi / 2)
    cases["uniform"] = (dv, 3, 2, c, n, a, m, xid)
    c, m, a, n, xid = 6, 3, 5, 4096, 3
    dv = np.zeros((c, m, a, 4), np.float32)
    dv[..., 0] = (rng.uniform(-16, 16, (c, m, a)) * TS).astype(np.float32)
    dv[..., 2] = rng.uniform(-np.pi, np.pi, (c, m, a)).astype(np.float32)
    dv[..., 1] = rng.standard_normal((c, m, a)).astype(np.float32)
    dv[..., 3] = rng.standard_normal((c, m, a)).astype(np.float32)
    cases["random"] = (dv, 1, 2, c, n, a, m, xid)
    for tag, (dv, b, p, c, n, a, m, xid) in cases.items():
        co64 = ref_coeffs(dv.astype(np.float64), b, p, c, n, a, m, xid)
        co32 = ref_coeffs(dv, b, p, c, n, a, m, xid)
        np.savez_compressed(
            os.path.join(HERE, f"golden_coeffs_{tag}.npz"),
            delay_vals=dv, coeffs_f64=co64, coeffs_f32=co32,
            params=np.array([b, p, c, n, a, m, xid], dtype=np.int64),
        )

    # ---- contraction: reference checker on the reference's own (uniform-delay) inputs ----
    b, a, c, t, m, n, xid = 2, 4, 3, 32, 2, 1024, 0
    x = rng.uniform(0, 255, (b, a, c, t, 2, 2)).astype(np.uint8)  # beamform_op_sequence_test.py:145-149
    dv = np.zeros((c, m, a, 4), np.float32)
    dv[..., 0] = np.single(5 * TS)
    dv[..., 2] = np.single(np.pi / 2)
    co = ref_coeffs(dv.astype(np.float64), b, 2, c, n, a, m, xid)
    re = ref_reorder.reorder(x, x.shape, (b, 2, c, t // 16, 16, a, 2))
    out = ref_mult.complex_mult(re, co, (b, 2, c, t // 16, 16, 2 * m))
    np.savez_compressed(
        os.path.join(HERE, "golden_pipeline_uniform.npz"),
        samples=x, delay_vals=dv, coeffs=co, reordered=re, beams=out,
        params=np.array([b, a, c, t, m, n, xid], dtype=np.int64),
    )

    # ---- contraction with NON-uniform coefficients: checker output kept to pin the beam-0 quirk ----
    b, a, c, t, m, n, xid = 1, 3, 2, 16, 3, 256, 1
    x = rng.integers(0, 256, (b, a, c, t, 2, 2), dtype=np.uint8)
    dv = np.zeros((c, m, a, 4), np.float32)
    dv[..., 0] = (rng.uniform(-16, 16, (c, m, a)) * TS).astype(np.float32)
    dv[..., 2] = rng.uniform(-np.pi, np.pi, (c, m, a)).astype(np.float32)
    co = ref_coeffs(dv.astype(np.float64), b, 2, c, n, a, m, xid)
    re = ref_reorder.reorder(x, x.shape, (b, 2, c, t // 16, 16, a, 2))
    out = ref_mult.complex_mult(re, co, (b, 2, c, t // 16, 16, 2 * m))
    np.savez_compressed(
        os.path.join(HERE, "golden_pipeline_random.npz"),
        samples=x, delay_vals=dv, coeffs=co, reordered=re, beams_beam0_checker=out,
        params=np.array([b, a, c, t, m, n, xid], dtype=np.int64),
    )
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
