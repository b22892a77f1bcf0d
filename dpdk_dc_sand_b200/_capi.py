"""ctypes binding of libdcbf.so (include/dcbf.h).  The ONLY compute path of this package.

There is deliberately no fallback: if the shared library is missing it is built with nvcc
(``dpdk_dc_sand_b200.build``); if that fails, or a call returns a non-zero status, an exception
is raised.  All pointers handed to the library are raw device (or, for the host plan, host)
addresses; PyTorch is only the owner of the memory behind them.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

OK = 0
ERR_INVALID_ARG = -1
ERR_UNSUPPORTED = -2
ERR_CUDA = -3
ERR_NO_DEVICE = -4
ERR_TIMEOUT = -5

FLAG_SIGNED_INPUT = 0x1
FLAG_FP16_COEFF = 0x2
FLAG_STREAMING = 0x4
FLAG_DEBUG_DIRECT_EPILOGUE = 0x100
FLAG_DEBUG_NO_KSTREAM = 0x200
FLAG_DEBUG_CUDA_CORES = 0x400
FLAG_DEBUG_WHOLE_CHANNELS = 0x800
FLAG_DEBUG_NO_PDL = 0x1000
FLAG_DEBUG_NO_PAIR = 0x2000
FLAG_DEBUG_NO_BEAM_PIECES = 0x4000
FLAG_DEBUG_TWO_A_STAGES = 0x8000

_ROLE_NAMES = {1: "producer", 2: "mma", 3: "epilogue", 4: "convert", 5: "coeff"}

# name -> (restype, argtypes); must list every symbol include/dcbf.h declares (tests check this).
SIGNATURES = {
    "dcbf_version": (C.c_int, []),
    "dcbf_strerror": (C.c_char_p, [C.c_int]),
    "dcbf_last_cuda_error": (C.c_char_p, []),
    "dcbf_reorder": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dcbf_coeffs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_double, C.c_void_p]),
    "dcbf_coeffs_tv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_double, C.POINTER(C.c_double), C.c_void_p]),
    "dcbf_beamform": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_uint, C.c_void_p]),
    "dcbf_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_int, C.c_double, C.c_uint, C.c_void_p]),
    "dcbf_fused_tv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_double, C.POINTER(C.c_double), C.c_uint, C.c_void_p]),
    "dcbf_fused_q8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.c_uint,
                                C.c_void_p]),
    "dcbf_fused_q8_bytes": (C.c_ulonglong, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "dcbf_fused_packed_bytes": (C.c_ulonglong, [C.c_int, C.c_int, C.c_int, C.c_uint]),
    "dcbf_fused_pack_coeffs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                         C.c_void_p, C.c_uint, C.c_void_p]),
    "dcbf_fused_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_double, C.c_uint, C.c_void_p]),
    "dcbf_fused_pack_coeffs_q8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_double, C.c_uint, C.c_void_p]),
    "dcbf_fused_packed_q8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint, C.c_void_p]),
    "dcbf_fused_status": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dcbf_fused_status_poll": (C.c_int, []),
    "dcbf_debug_set_profile_buffer": (None, [C.c_void_p]),
    "dcbf_fused_tiling": (None, [C.c_int, C.c_int, C.c_uint, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int)]),
    "dcbf_host_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_double, C.c_uint, C.c_int, C.c_int]),
    "dcbf_host_plan_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcbf_host_plan_set_delay_vals": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dcbf_host_plan_set_gains": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dcbf_host_plan_run_q8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_ulonglong)]),
    "dcbf_host_plan_destroy": (C.c_int, [C.c_void_p]),
    "dcbf_launch_count": (C.c_ulonglong, []),
    "dcbf_fused_bytes": (C.c_ulonglong, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
}

class FusedOptions(C.Structure):
    """dcbf_fused_options (include/dcbf.h)."""

    _fields_ = [("struct_size", C.c_size_t), ("batch_dt_s", C.POINTER(C.c_double)), ("beam_weights", C.c_void_p),
                ("beam_gains", C.c_void_p), ("beams_q8", C.c_void_p), ("saturated", C.c_void_p),
                ("sample_dt_s", C.c_double), ("beam_weights_log2", C.c_int)]


SIGNATURES["dcbf_fused_ex"] = (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_double, C.POINTER(FusedOptions), C.c_uint, C.c_void_p])
SIGNATURES["dcbf_coeffs_f16"] = (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_double, C.POINTER(C.c_double), C.c_void_p, C.c_void_p])
SIGNATURES["dcbf_coeffs_ex"] = (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_double, C.POINTER(C.c_double), C.c_void_p, C.c_void_p])

SIGNATURES.update({
    "dcbf_ingest_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong,
                                     C.c_int]),
    "dcbf_ingest_heap": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "dcbf_ingest_heap_ptr": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.POINTER(C.c_void_p)]),
    "dcbf_ingest_heap_done": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int]),
    "dcbf_ingest_packet": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "dcbf_ingest_set_frequency": (C.c_int, [C.c_void_p, C.c_longlong]),
    "dcbf_ingest_pop": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong),
                                  C.POINTER(C.c_int), C.c_void_p]),
    "dcbf_ingest_release": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dcbf_ingest_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong),
                                    C.POINTER(C.c_ulonglong)]),
    "dcbf_ingest_destroy": (C.c_int, [C.c_void_p]),
})

_lib = None
_lock = threading.Lock()


class DcbfError(RuntimeError):
    """A libdcbf call failed (CUDA error, unsupported shape, watchdog)."""


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """Load (building first if needed) libdcbf.so and attach prototypes.  Raises if impossible."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("DCBF_LIB") or _build.LIB_PATH  # DCBF_LIB: developer A/B runs against another build
        if not os.path.exists(path):
            path = _build.build()
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def strerror(status: int) -> str:
    return load().dcbf_strerror(status).decode()


def check(status: int, what: str) -> None:
    if status == OK:
        return
    lib = load()
    msg = f"{what}: {lib.dcbf_strerror(status).decode()} ({status})"
    if status == ERR_CUDA:
        msg += f" [{lib.dcbf_last_cuda_error().decode()}]"
    if status in (ERR_INVALID_ARG, ERR_UNSUPPORTED):
        raise ValueError(msg)
    raise DcbfError(msg)


def _ptr(t) -> int:
    """Device address of a torch tensor (must be contiguous)."""
    if not t.is_contiguous():
        raise ValueError("libdcbf needs contiguous buffers")
    return t.data_ptr()


def _stream_handle(stream) -> int:
    if stream is None:
        return 0
    return int(getattr(stream, "cuda_stream", stream))


def reorder(samples, reordered, n_batches, n_ants, n_chans, n_samples, stream=None) -> None:
    check(load().dcbf_reorder(_ptr(samples), _ptr(reordered), n_batches, n_ants, n_chans, n_samples,
                              _stream_handle(stream)), "dcbf_reorder")


def _dt_array(batch_dt, n_batches):
    if len(batch_dt) != n_batches:
        raise ValueError("batch_dt needs one entry (seconds) per batch")
    return (C.c_double * n_batches)(*[float(t) for t in batch_dt])


def coeffs(delay_vals, out, n_batches, n_pols, n_chans, n_chans_total, n_ants, n_beams, xeng_id, sample_period,
           stream=None, batch_dt=None, weights=None) -> None:
    """Stand-alone steering coefficients; a float16 ``out`` tensor selects dcbf_coeffs_f16 (same layout, fp16)."""
    if getattr(out, "dtype", None) is not None and str(out.dtype).endswith("float16"):
        dt = _dt_array(batch_dt, n_batches) if batch_dt is not None else None
        check(load().dcbf_coeffs_f16(_ptr(delay_vals), _ptr(out), n_batches, n_pols, n_chans, n_chans_total, n_ants,
                                     n_beams, xeng_id, float(sample_period), dt,
                                     _ptr(weights) if weights is not None else None, _stream_handle(stream)),
              "dcbf_coeffs_f16")
        return
    if weights is not None:
        dt = _dt_array(batch_dt, n_batches) if batch_dt is not None else None
        check(load().dcbf_coeffs_ex(_ptr(delay_vals), _ptr(out), n_batches, n_pols, n_chans, n_chans_total, n_ants,
                                    n_beams, xeng_id, float(sample_period), dt, _ptr(weights), _stream_handle(stream)),
              "dcbf_coeffs_ex")
        return
    if batch_dt is not None:
        check(load().dcbf_coeffs_tv(_ptr(delay_vals), _ptr(out), n_batches, n_pols, n_chans, n_chans_total, n_ants,
                                    n_beams, xeng_id, float(sample_period), _dt_array(batch_dt, n_batches),
                                    _stream_handle(stream)), "dcbf_coeffs_tv")
        return
    check(load().dcbf_coeffs(_ptr(delay_vals), _ptr(out), n_batches, n_pols, n_chans, n_chans_total, n_ants, n_beams,
                             xeng_id, float(sample_period), _stream_handle(stream)), "dcbf_coeffs")


def beamform(reordered, coeff, beams, n_batches, n_chans, n_samples, n_ants, n_beams, flags=0, stream=None) -> None:
    check(load().dcbf_beamform(_ptr(reordered), _ptr(coeff), _ptr(beams), n_batches, n_chans, n_samples, n_ants,
                               n_beams, flags, _stream_handle(stream)), "dcbf_beamform")


def fused_ex(samples, delay_vals, beams, n_batches, n_ants, n_chans, n_chans_total, n_samples, n_beams, xeng_id,
             sample_period, flags=0, stream=None, batch_dt=None, weights=None, gains=None, beams_q8=None,
             saturated=None, sample_dt=0.0, weights_log2=0) -> None:
    """dcbf_fused_ex: any combination of per-heap (and, with ``sample_dt``, per-time-tile) times, per-(beam, antenna)
    weights and int8 output."""
    opts = FusedOptions()
    opts.struct_size = C.sizeof(FusedOptions)
    dt = _dt_array(batch_dt, n_batches) if batch_dt is not None else None  # keep alive until the call returns
    if dt is not None:
        opts.batch_dt_s = C.cast(dt, C.POINTER(C.c_double))
    opts.beam_weights = _ptr(weights) if weights is not None else None
    opts.beam_gains = _ptr(gains) if gains is not None else None
    opts.beams_q8 = _ptr(beams_q8) if beams_q8 is not None else None
    opts.saturated = _ptr(saturated) if saturated is not None else None
    opts.sample_dt_s = float(sample_dt or 0.0)
    opts.beam_weights_log2 = int(weights_log2) if weights is not None else 0
    check(load().dcbf_fused_ex(_ptr(samples), _ptr(delay_vals), _ptr(beams) if beams is not None else None, n_batches,
                               n_ants, n_chans, n_chans_total, n_samples, n_beams, xeng_id, float(sample_period),
                               C.byref(opts), flags, _stream_handle(stream)), "dcbf_fused_ex")


def fused(samples, delay_vals, beams, n_batches, n_ants, n_chans, n_chans_total, n_samples, n_beams, xeng_id,
          sample_period, flags=0, stream=None, batch_dt=None, weights=None, sample_dt=0.0, weights_log2=0) -> None:
    if weights is not None or sample_dt:
        fused_ex(samples, delay_vals, beams, n_batches, n_ants, n_chans, n_chans_total, n_samples, n_beams, xeng_id,
                 sample_period, flags, stream, batch_dt=batch_dt, weights=weights, sample_dt=sample_dt,
                 weights_log2=weights_log2)
        return
    if batch_dt is not None:
        check(load().dcbf_fused_tv(_ptr(samples), _ptr(delay_vals), _ptr(beams), n_batches, n_ants, n_chans,
                                   n_chans_total, n_samples, n_beams, xeng_id, float(sample_period),
                                   _dt_array(batch_dt, n_batches), flags, _stream_handle(stream)), "dcbf_fused_tv")
        return
    check(load().dcbf_fused(_ptr(samples), _ptr(delay_vals), _ptr(beams), n_batches, n_ants, n_chans, n_chans_total,
                            n_samples, n_beams, xeng_id, float(sample_period), flags, _stream_handle(stream)),
          "dcbf_fused")


def fused_q8(samples, delay_vals, gains, beams_q8, n_batches, n_ants, n_chans, n_chans_total, n_samples, n_beams,
             xeng_id, sample_period, flags=0, stream=None, batch_dt=None, saturated=None) -> None:
    """Fused path with int8 requantised output; ``saturated`` is an optional 1-element int64 CUDA tensor (counter)."""
    dt = _dt_array(batch_dt, n_batches) if batch_dt is not None else None
    check(load().dcbf_fused_q8(_ptr(samples), _ptr(delay_vals), _ptr(gains), _ptr(beams_q8),
                               _ptr(saturated) if saturated is not None else None, n_batches, n_ants, n_chans,
                               n_chans_total, n_samples, n_beams, xeng_id, float(sample_period), dt, flags,
                               _stream_handle(stream)), "dcbf_fused_q8")


def fused_q8_bytes(n_batches, n_ants, n_chans, n_samples, n_beams) -> int:
    return int(load().dcbf_fused_q8_bytes(n_batches, n_ants, n_chans, n_samples, n_beams))


def fused_status() -> None:
    """Synchronise and raise if a fused kernel's pipeline watchdog fired."""
    role, barrier, block = C.c_int(0), C.c_int(0), C.c_int(0)
    st = load().dcbf_fused_status(C.byref(role), C.byref(barrier), C.byref(block))
    if st == ERR_TIMEOUT:
        raise DcbfError(f"dcbf_fused watchdog: role {_ROLE_NAMES.get(role.value, role.value)} stuck on barrier "
                        f"{barrier.value} in block {block.value}")
    check(st, "dcbf_fused_status")


def fused_status_poll() -> None:
    """Cheap form for work the caller has already synchronised with (no CUDA call unless a kernel raised its flag)."""
    if _lib is not None and _lib.dcbf_fused_status_poll() != OK:
        fused_status()


def fused_tiling(n_ants, n_beams, flags=0):
    kb, nt, ntc = C.c_int(0), C.c_int(0), C.c_int(0)
    load().dcbf_fused_tiling(n_ants, n_beams, flags, C.byref(kb), C.byref(nt), C.byref(ntc))
    return kb.value, nt.value, ntc.value


def fused_packed_bytes(n_ants, n_chans, n_beams, flags=0) -> int:
    """Bytes of the packed steering-coefficient tile sets of n_chans channels (0: the shape keeps no whole tile set)."""
    return int(load().dcbf_fused_packed_bytes(n_ants, n_chans, n_beams, flags))


def fused_pack_coeffs(delay_vals, packed, n_ants, n_chans, n_chans_total, n_beams, xeng_id, sample_period, flags=0,
                      stream=None, weights=None) -> None:
    """dcbf_fused_pack_coeffs: one delay model -> tile sets in the tensor cores' layout (device uint8 tensor `packed`)."""
    check(load().dcbf_fused_pack_coeffs(_ptr(delay_vals), _ptr(packed), n_ants, n_chans, n_chans_total, n_beams, xeng_id,
                                        float(sample_period), _ptr(weights) if weights is not None else None, flags,
                                        _stream_handle(stream)), "dcbf_fused_pack_coeffs")


def fused_packed(samples, packed, beams, n_batches, n_ants, n_chans, n_chans_total, n_samples, n_beams, xeng_id,
                 sample_period, flags=0, stream=None) -> None:
    """dcbf_fused_packed: dcbf_fused with the coefficients read from `packed` instead of being evaluated per heap."""
    check(load().dcbf_fused_packed(_ptr(samples), _ptr(packed), _ptr(beams), n_batches, n_ants, n_chans, n_chans_total,
                                   n_samples, n_beams, xeng_id, float(sample_period), flags, _stream_handle(stream)),
          "dcbf_fused_packed")


def fused_pack_coeffs_q8(delay_vals, gains, packed, n_ants, n_chans, n_chans_total, n_beams, xeng_id, sample_period,
                         flags=0, stream=None) -> None:
    """dcbf_fused_pack_coeffs_q8: tile sets for fused_packed_q8 (the quantisation gains are folded in)."""
    check(load().dcbf_fused_pack_coeffs_q8(_ptr(delay_vals), _ptr(gains), _ptr(packed), n_ants, n_chans, n_chans_total,
                                           n_beams, xeng_id, float(sample_period), flags, _stream_handle(stream)),
          "dcbf_fused_pack_coeffs_q8")


def fused_packed_q8(samples, packed, gains, beams_q8, n_batches, n_ants, n_chans, n_chans_total, n_samples, n_beams,
                    xeng_id, sample_period, flags=0, stream=None, saturated=None) -> None:
    """dcbf_fused_packed_q8: fused_q8 on packed coefficients."""
    check(load().dcbf_fused_packed_q8(_ptr(samples), _ptr(packed), _ptr(gains), _ptr(beams_q8),
                                      _ptr(saturated) if saturated is not None else None, n_batches, n_ants, n_chans,
                                      n_chans_total, n_samples, n_beams, xeng_id, float(sample_period), flags,
                                      _stream_handle(stream)), "dcbf_fused_packed_q8")


def fused_bytes(n_batches, n_ants, n_chans, n_samples, n_beams) -> int:
    return int(load().dcbf_fused_bytes(n_batches, n_ants, n_chans, n_samples, n_beams))


def launch_count() -> int:
    return int(load().dcbf_launch_count())


class HostPlan:
    """dcbf_host_plan_*: fused path on HOST arrays (numpy / pinned), chunked + pipelined over PCIe."""

    def __init__(self, n_batches, n_ants, n_chans, n_chans_total, n_samples, n_beams, xeng_id, sample_period,
                 flags=0, chunk_chans=0, n_slots=3):
        self._h = C.c_void_p(None)
        self.shape_in = (n_batches, n_ants, n_chans, n_samples, 2, 2)
        self.shape_dv = (n_chans, n_beams, n_ants, 4)
        self.shape_out = (n_batches, 2, n_chans, n_samples // 16, 16, 2 * n_beams)
        check(load().dcbf_host_plan_create(C.byref(self._h), n_batches, n_ants, n_chans, n_chans_total, n_samples,
                                           n_beams, xeng_id, float(sample_period), flags, chunk_chans, n_slots),
              "dcbf_host_plan_create")

    @staticmethod
    def _check_arrays(triples) -> None:
        import numpy as np

        for arr, shape, dt in triples:
            if arr is None:
                continue
            if tuple(arr.shape) != shape or arr.dtype != dt or not arr.flags["C_CONTIGUOUS"]:
                raise ValueError(f"expected C-contiguous {np.dtype(dt).name} array of shape {shape}, got "
                                 f"{arr.dtype} {arr.shape}")

    def set_delay_vals(self, delay_vals) -> None:
        """Upload the delay model once; ``run(samples, None, beams)`` then moves only the voltages per step."""
        import numpy as np

        self._check_arrays([(delay_vals, self.shape_dv, np.float32)])
        check(load().dcbf_host_plan_set_delay_vals(self._h, delay_vals.ctypes.data), "dcbf_host_plan_set_delay_vals")

    def run(self, samples, delay_vals, beams) -> None:
        """numpy arrays (C-contiguous; pinned for overlap) with the reference shapes; blocks until done.
        ``delay_vals=None``: the resident delay model of ``set_delay_vals``."""
        import numpy as np

        self._check_arrays([(samples, self.shape_in, np.uint8), (delay_vals, self.shape_dv, np.float32),
                            (beams, self.shape_out, np.float32)])
        check(load().dcbf_host_plan_run(self._h, samples.ctypes.data,
                                        delay_vals.ctypes.data if delay_vals is not None else None, beams.ctypes.data),
              "dcbf_host_plan_run")

    def set_gains(self, gains) -> None:
        """Per-beam quantisation gains (float32 host array of n_beams) for ``run_q8``."""
        import numpy as np

        g = np.ascontiguousarray(gains, dtype=np.float32)
        if g.shape != (self.shape_dv[1],):
            raise ValueError("gains must have one entry per beam")
        check(load().dcbf_host_plan_set_gains(self._h, g.ctypes.data), "dcbf_host_plan_set_gains")

    def run_q8(self, samples, delay_vals, beams_q8) -> int:
        """Like ``run`` with int8 requantised beams; returns the number of clipped values."""
        import numpy as np

        self._check_arrays([(samples, self.shape_in, np.uint8), (delay_vals, self.shape_dv, np.float32),
                            (beams_q8, self.shape_out, np.int8)])
        sat = C.c_ulonglong(0)
        check(load().dcbf_host_plan_run_q8(self._h, samples.ctypes.data,
                                           delay_vals.ctypes.data if delay_vals is not None else None,
                                           beams_q8.ctypes.data, C.byref(sat)), "dcbf_host_plan_run_q8")
        return int(sat.value)

    def close(self) -> None:
        if self._h:
            load().dcbf_host_plan_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Ingest:
    """dcbf_ingest_*: assembles F-engine heaps into ``(n_batches, n_ants, n_chans, n_samples, 2, 2)`` chunks."""

    def __init__(self, n_chunks, n_batches, n_ants, n_chans, n_samples, timestamp_step, pinned=True):
        self._h = C.c_void_p(None)
        self.shape = (n_batches, n_ants, n_chans, n_samples, 2, 2)
        self.heap_shape = (n_chans, n_samples, 2, 2)
        check(load().dcbf_ingest_create(C.byref(self._h), n_chunks, n_batches, n_ants, n_chans, n_samples,
                                        int(timestamp_step), 1 if pinned else 0), "dcbf_ingest_create")

    def heap(self, timestamp, feng_id, payload) -> bool:
        """Place one heap (any 8-bit array of ``heap_shape``); False if it was dropped (late / no free chunk)."""
        import numpy as np

        arr = np.ascontiguousarray(payload)
        if arr.dtype.itemsize != 1 or tuple(arr.shape) != self.heap_shape:
            raise ValueError(f"heap payload must be an 8-bit array of shape {self.heap_shape}")
        st = load().dcbf_ingest_heap(self._h, int(timestamp), int(feng_id), arr.ctypes.data)
        if st == ERR_UNSUPPORTED:
            return False
        check(st, "dcbf_ingest_heap")
        return True

    def packet(self, data, default_feng_id=-1) -> bool:
        """Place the payload of one raw SPEAD-64-48 packet (bytes-like); False if it was dropped or carries no data
        (late, another sub-band, descriptor heap); ValueError for a malformed packet."""
        buf = bytes(data)
        st = load().dcbf_ingest_packet(self._h, buf, len(buf), int(default_feng_id))
        if st == ERR_UNSUPPORTED:
            return False
        check(st, "dcbf_ingest_packet")
        return True

    def set_frequency(self, first_channel) -> None:
        """Only accept heaps whose frequency item (0x4103) equals ``first_channel`` (-1: any)."""
        check(load().dcbf_ingest_set_frequency(self._h, int(first_channel)), "dcbf_ingest_set_frequency")

    def pop(self, flush=False):
        """Next finished chunk as ``(samples view, first_timestamp, present[B, A])`` or None.  The view aliases the
        ring's memory: call ``release(samples)`` when done with it."""
        import numpy as np

        ptr, ts, missing = C.c_void_p(None), C.c_longlong(0), C.c_int(0)
        present = np.zeros(self.shape[:2], np.uint8)
        got = load().dcbf_ingest_pop(self._h, 1 if flush else 0, C.byref(ptr), C.byref(ts), C.byref(missing),
                                     present.ctypes.data)
        if got < 0:
            check(got, "dcbf_ingest_pop")
        if got == 0:
            return None
        n = int(np.prod(self.shape))
        buf = (C.c_uint8 * n).from_address(ptr.value)
        samples = np.frombuffer(buf, dtype=np.uint8).reshape(self.shape)
        assert int(missing.value) == int(present.size - present.sum())
        return samples, int(ts.value), present.astype(bool)

    def release(self, samples) -> None:
        check(load().dcbf_ingest_release(self._h, samples.ctypes.data), "dcbf_ingest_release")

    def stats(self) -> dict:
        a, b, c = C.c_ulonglong(0), C.c_ulonglong(0), C.c_ulonglong(0)
        check(load().dcbf_ingest_stats(self._h, C.byref(a), C.byref(b), C.byref(c)), "dcbf_ingest_stats")
        return {"late_or_dropped": int(a.value), "duplicate": int(b.value), "bad": int(c.value)}

    def close(self) -> None:
        if self._h:
            load().dcbf_ingest_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
