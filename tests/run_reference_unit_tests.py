"""Run the reference's OWN unit-test files, unmodified, against the drop-in.

    python tests/run_reference_unit_tests.py <path to the reference's beamformer/unit_test> [extra pytest arguments]

The four files (beamform_op_sequence_test.py, prebeamform_reorder_test.py, beamform_coeff_test.py,
beamform_mult_kernel_test.py; reference: beamformer/unit_test/) import `beamforming.*` and `katsdpsigproc.accel`
and compare the GPU operators with the reference's own CPU checkers (unit_test/coeff_generator_cpu.py,
complex_mult_cpu.py, beamforming/reorder.py).  Here `install_dropin()` makes those imports resolve to this package --
`beamforming` is dpdk_dc_sand_b200.beamforming (C ABI -> libdcbf.so), `katsdpsigproc` the torch-backed runtime shim,
the `katsdpsigproc.pytest_plugin` named by the reference's conftest.py:41 included -- and pytest runs the files as they
are, with `--all-combinations` (conftest.py:61-101).  The CPU checkers stay the reference's own files.

Nothing of the reference is part of this repository: stage the directory next to the run (the GPU box has no
/root/reference), e.g. `cp -r /root/reference/beamformer/unit_test gpurun_in/unit_test` (gpurun_in/ is git-ignored).
Two environment shims, both for running numpy-1.x-era code under numpy >= 2, neither touching a test file:
  * `np.math = math` (numpy 2 removed `np.math`, which coeff_generator_cpu.py:148 uses);
  * the CPU checker's `delay_vals` are widened to float64 on entry (CoeffGenerator.__init__): under numpy 1.x a float32
    SCALAR times a Python number was float64 (value-based promotion), so coeff_generator_cpu.py:143-165 evaluated the
    phase in float64; under numpy >= 2 (NEP 50) the same source evaluates in float32 and its exact-equality assertion
    (beamform_coeff_test.py:172) can then only be met by a float32 phase, which the reference's GPU kernel does not
    compute either.  float32 -> float64 is exact, so this reproduces the numpy-1.x result bit for bit
    (tests/golden/make_golden.py keeps both evaluations).  `--numpy2-checker` runs without this shim.
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    if len(sys.argv) < 2 or not os.path.isdir(sys.argv[1]):
        sys.stderr.write(__doc__)
        return 2
    unit_test = os.path.abspath(sys.argv[1])
    import numpy as np
    import pytest

    if not hasattr(np, "math"):
        np.math = math
    import dpdk_dc_sand_b200

    dpdk_dc_sand_b200.install_dropin(force_accel_shim=True)
    sys.path.insert(0, os.path.dirname(unit_test))  # `from unit_test import ...`
    extra = list(sys.argv[2:])
    if "--numpy2-checker" in extra:
        extra.remove("--numpy2-checker")
    else:
        from unit_test import coeff_generator_cpu

        plain_init = coeff_generator_cpu.CoeffGenerator.__init__

        def init_numpy1(self, delay_vals, *a, **k):
            plain_init(self, np.asarray(delay_vals, dtype=np.float64), *a, **k)

        coeff_generator_cpu.CoeffGenerator.__init__ = init_numpy1
    sys.argv[2:] = extra
    args = [unit_test, "--all-combinations", "-q", "-p", "no:cacheprovider", "--rootdir", unit_test,
            "-o", "python_files=*_test.py", *sys.argv[2:]]
    return int(pytest.main(args))


if __name__ == "__main__":
    sys.exit(main())
