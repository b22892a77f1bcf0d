"""dcbf: B200-native (sm_100a) tied-array beamforming hot path behind dc_sand's ``beamforming`` operator API.

Layout
    csrc/            CUDA kernels + the C ABI (include/dcbf.h) -> lib/libdcbf.so
    _capi.py         ctypes binding of that library (no other compute path exists)
    katsdpsigproc/   minimal katsdpsigproc.accel-compatible runtime over torch CUDA buffers/streams
    beamforming/     the reference's operator classes (same names, ctor args, slots, shapes)
    sharding.py      frequency-channel sharding across GPUs (rank == xeng_id) + optional NCCL beam gather

``install_dropin()`` aliases ``beamforming`` and (when the real one is absent) ``katsdpsigproc`` in
``sys.modules`` so code written against the reference imports runs unchanged.
"""
from __future__ import annotations

import importlib
import sys

__version__ = "0.1.0"


def install_dropin(force_accel_shim: bool = False) -> None:
    """Make ``import beamforming`` / ``from katsdpsigproc import accel`` resolve to this package."""
    have_real = False
    if not force_accel_shim:
        try:
            mod = importlib.import_module("katsdpsigproc")
            have_real = not getattr(mod, "__dcbf_shim__", False)
        except ImportError:
            have_real = False
    if not have_real:
        shim = importlib.import_module("dpdk_dc_sand_b200.katsdpsigproc")
        sys.modules["katsdpsigproc"] = shim
        for sub in ("accel", "abc", "pytest_plugin"):
            sys.modules[f"katsdpsigproc.{sub}"] = importlib.import_module(f"dpdk_dc_sand_b200.katsdpsigproc.{sub}")
    bf = importlib.import_module("dpdk_dc_sand_b200.beamforming")
    sys.modules["beamforming"] = bf
    for sub in ("prebeamform_reorder", "coeff_generator", "matrix_multiply", "complex_mult_kernel",
                "beamform_op_sequence", "reorder"):
        sys.modules[f"beamforming.{sub}"] = importlib.import_module(f"dpdk_dc_sand_b200.beamforming.{sub}")
