"""Authoring-container probe (needs /root/reference, so it cannot run on the GPU box): the reference's OWN CPU functions
of the path, unmodified, timed on this container's cores at C1 and on a 1/16 sample of C2.

    python tools/time_verbatim_reference.py   ->  profiles/r02_cpu_verbatim_reference.json

    beamformer/beamforming/reorder.py:reorder                  (numba)
    beamformer/unit_test/coeff_generator_cpu.py:cpu_coeffs     (pure python loops)
    beamformer/unit_test/complex_mult_cpu.py:complex_mult      (numba; beam-0 shortcut of the reference checker)

bench.py reports the oracle's vectorised numpy port as `cpu_baseline` (it can travel to the GPU box); this file is the
context for how much faster that port is than the code it restates.
"""
import json
import math
import os
import sys
import time

import numpy as np

np.math = math  # numpy >= 2 shim (coeff_generator_cpu.py:148)
sys.path.insert(0, "/root/reference/beamformer")
from beamforming import reorder as ref_reorder  # noqa: E402
from unit_test import complex_mult_cpu as ref_mult  # noqa: E402
from unit_test.coeff_generator_cpu import CoeffGenerator as RefCoeffGenerator  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TS = 1 / 1712e6


def time_case(name, b, a, c, t, m, n_total, coeff_chans):
    rng = np.random.default_rng(5)
    x = rng.integers(0, 256, (b, a, c, t, 2, 2), dtype=np.uint8)
    dv = np.zeros((c, m, a, 4), np.float32)
    dv[..., 0] = (rng.uniform(-16, 16, (c, m, a)) * TS).astype(np.float32)
    dv[..., 2] = rng.uniform(-np.pi, np.pi, (c, m, a)).astype(np.float32)
    out_shape = (b, 2, c, t // 16, 16, a, 2)
    x1 = np.ascontiguousarray(x[:, :, :1])
    ref_reorder.reorder(x1, x1.shape, (b, 2, 1, t // 16, 16, a, 2))  # numba compile
    t0 = time.perf_counter()
    re = ref_reorder.reorder(x, x.shape, out_shape)
    t_reorder = time.perf_counter() - t0
    # cpu_coeffs is a pure-python triple loop: time `coeff_chans` channels and scale
    cc = min(c, coeff_chans)
    t0 = time.perf_counter()
    co_part = RefCoeffGenerator(dv[:cc].astype(np.float64), b, 2, cc, n_total, 16, 16, a, m, 0, TS).cpu_coeffs()
    t_coeff = (time.perf_counter() - t0) * c / cc
    co = np.zeros((b, 2, c, 2 * a, 2 * m), np.float32)
    co[:, :, :cc] = co_part
    ref_mult.complex_mult(np.ascontiguousarray(re[:, :, :1]), np.ascontiguousarray(co[:, :, :1]), (b, 2, 1, t // 16, 16, 2 * m))  # numba compile
    t0 = time.perf_counter()
    ref_mult.complex_mult(re, co, (b, 2, c, t // 16, 16, 2 * m))
    t_mult = time.perf_counter() - t0
    total = t_reorder + t_coeff + t_mult
    return {"case": name, "n_batches": b, "n_ants": a, "n_chans": c, "n_samples": t, "n_beams": m,
            "reorder_s": t_reorder, "cpu_coeffs_s": t_coeff, "cpu_coeffs_timed_channels": cc, "complex_mult_s": t_mult,
            "total_s": total, "input_GBps": x.nbytes / total / 1e9}


def main():
    cases = [time_case("c1: 4 antennas x 64 channels x 256 samples, 4 beams", 1, 4, 64, 256, 4, 64, 64),
             time_case("1/16 of c2: 64 antennas x 64 of 1024 channels x 256 samples, 16 beams", 1, 64, 64, 256, 16, 1024, 4)]
    out = {"what": "the reference's own CPU functions (reorder.reorder, CoeffGenerator.cpu_coeffs, complex_mult), unmodified, "
                   "in the authoring container; complex_mult is the reference checker's beam-0 shortcut",
           "cores": os.cpu_count(), "python": sys.version.split()[0], "numpy": np.__version__, "cases": cases}
    path = os.path.join(ROOT, "profiles", "r02_cpu_verbatim_reference.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
