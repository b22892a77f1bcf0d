"""Steering-coefficient generation operation (stand-alone stage 2).

API mirror of ``beamformer/beamforming/coeff_generator.py`` (template :106-181, operation :184-250).  The
numba kernel ``run_coeff_gen`` (:12-103) is replaced by ``dcbf_coeffs`` (csrc/coeffs.cu): float64 phase in the
reference's operation order, float64 sincos, float32 store; ``delay_vals[c][beam][ant]`` lands at rows
``2*ant``, columns ``2*beam`` (the consistent indexing of ``unit_test/coeff_generator_cpu.py:125-186``; the
reference GPU kernel's read index transposes ant/beam, which only uniform delays hide).  The launch does not
host-synchronise (the reference calls ``cuda.synchronize()``, :250).
"""
import math

import numpy as np

from .. import _capi
from ..katsdpsigproc import accel
from ..katsdpsigproc.accel import IOSlot, Operation


class CoeffGeneratorTemplate:
    """Same constructor as the reference (coeff_generator.py:138-151)."""

    def __init__(self, context, n_batches: int, n_pols: int, n_channels_per_stream: int, n_channels: int,
                 n_blocks: int, n_samples_per_block: int, n_ants: int, n_beams: int, xeng_id: int,
                 sample_period: float) -> None:
        self.context = context
        self.n_batches = n_batches
        self.n_pols = n_pols
        self.n_channels_per_stream = n_channels_per_stream
        self.n_channels = n_channels
        self.n_blocks = n_blocks
        self.n_samples_per_block = n_samples_per_block
        self.n_ants = n_ants
        self.n_beams = n_beams
        self.xeng_id = xeng_id
        self.sample_period = sample_period

        # {delay_s, delay_rate, phase_rad, phase_rate} per (channel, beam, antenna)            coeff_generator.py:164-169
        self.delay_vals_data_dimensions = accel.exact_dimensions(n_channels_per_stream, n_beams, n_ants, 4)
        # one [[cos, sin], [-sin, cos]] block per (antenna, beam), replicated over batch and pol    :171-177
        self.coeff_data_dimensions = accel.exact_dimensions(n_batches, n_pols, n_channels_per_stream, 2 * n_ants,
                                                            2 * n_beams)

    def instantiate(self, command_queue) -> "CoeffGenerator":
        return CoeffGenerator(self, command_queue)


def _device_weights(op):
    """``op.beam_weights`` (None, a CUDA tensor, or anything array-like of shape (n_beams, n_ants)) as a device tensor."""
    w = op.beam_weights
    if w is None:
        return None
    import torch

    t = op.template
    n_beams = getattr(t, "n_beams", None) or op.template.beamform_coeff_template.n_beams
    n_ants = getattr(t, "n_ants", None) or op.template.beamform_coeff_template.n_ants
    if not (isinstance(w, torch.Tensor) and w.is_cuda):
        dev = op.buffer("delay_vals" if "delay_vals" in op.slots else "bufin_delay_vals").buffer.device
        w = torch.as_tensor(np.ascontiguousarray(w, dtype=np.float32)).to(dev)
        op.beam_weights = w  # uploaded once; assign a new array to change the weights
    if tuple(w.shape) != (n_beams, n_ants) or w.dtype != torch.float32:
        raise ValueError("beam_weights must be float32 of shape (n_beams, n_ants)")
    w = w.contiguous()
    # power-of-two bound of the weights for the fused kernel (dcbf_fused_options.beam_weights_log2), found once per
    # tensor: max|w| <= 2^e keeps the fp16 coefficient pair at full precision for weights from 6e-5 to 3e4
    if getattr(op, "_beam_weights_log2_of", None) is not w:
        wmax = float(w.abs().max().item()) if w.numel() else 0.0
        if not math.isfinite(wmax):
            raise ValueError("beam_weights must be finite")
        e = 0 if wmax == 0.0 else int(math.ceil(math.log2(wmax)))
        if e > 15:
            raise ValueError(f"beam_weights up to {wmax:g}: the fused path takes |w| <= 32768")
        op._beam_weights_log2 = max(e, -14)
        op._beam_weights_log2_of = w
        op.beam_weights = w
    return w


def weights_log2(op) -> int:
    """The power-of-two bound `_device_weights` found for ``op.beam_weights`` (0 without weights)."""
    return int(getattr(op, "_beam_weights_log2", 0)) if op.beam_weights is not None else 0


class CoeffGenerator(Operation):
    """.. rubric:: Slots

    delay_vals: (n_channels_per_stream, n_beams, n_ants, 4), float32 -- {delay_s, delay_rate, phase_rad, phase_rate}
    outCoeffs: (n_batches, n_pols, n_channels_per_stream, 2*n_ants, 2*n_beams), float32

    ``batch_times`` (attribute, default None = the reference behaviour: rates ignored): one time offset in seconds
    per batch, measured from the delay model's reference time; batch b is then steered with
    ``delay + delay_rate*t_b`` and ``phase + phase_rate*t_b`` (the native precursor's time-varying form,
    beamformer_coefficient_generator/BeamformerKernels.cu:25-35).
    ``beam_weights`` (attribute, default None): real weights of shape (n_beams, n_ants) multiplied into the
    coefficients -- the payload of the control plane's ``?beam-weights`` request (ngkcs/ngkcs/corr3_servlet.py:140).
    """

    def __init__(self, template: CoeffGeneratorTemplate, command_queue) -> None:
        super().__init__(command_queue)
        self.template = template
        self.batch_times = None
        self.beam_weights = None
        self.slots["delay_vals"] = IOSlot(dimensions=template.delay_vals_data_dimensions, dtype=np.float32)
        self.slots["outCoeffs"] = IOSlot(dimensions=template.coeff_data_dimensions, dtype=np.float32)

    def _run(self) -> None:
        t = self.template
        _capi.coeffs(self.buffer("delay_vals").buffer, self.buffer("outCoeffs").buffer, t.n_batches, t.n_pols,
                     t.n_channels_per_stream, t.n_channels, t.n_ants, t.n_beams, t.xeng_id, t.sample_period,
                     self.command_queue.stream, batch_dt=self.batch_times, weights=_device_weights(self))
