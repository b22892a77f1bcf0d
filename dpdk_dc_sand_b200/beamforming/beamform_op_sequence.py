"""Reorder + coefficient generation + beamform multiplication as one operation.

API mirror of ``beamformer/beamforming/beamform_op_sequence.py`` (template :16-114, sequence :117-157): same
constructor, same sub-operation attributes (``prebeamform_reorder``, ``beamform_coeff``, ``beamform_mult``),
same compound buffer names (``bufin_reorder``, ``bufin_delay_vals``, ``bufout_mult``, ``bufint_data``,
``bufint_coeff``).

Where the reference launches three kernels that round-trip the reordered voltages and the coefficients
through device memory, ``OpSequence._run`` launches ONE fused sm_100a kernel (``dcbf_fused``, csrc/fused.cu)
that reads every voltage byte and every delay value once.  The two intermediates are therefore not needed:
``ensure_all_bound()`` allocates only the three external buffers unless ``materialize_intermediates`` is set
(or the caller bound them), in which case the stand-alone reorder / coefficient kernels fill them as well so
that code inspecting ``bufint_*`` after a call still finds the reference's contents.
"""
from .. import _capi
from ..katsdpsigproc import accel
from .coeff_generator import CoeffGeneratorTemplate, _device_weights, weights_log2
from .matrix_multiply import MatrixMultiplyTemplate
from .prebeamform_reorder import PreBeamformReorderTemplate

_INTERMEDIATES = ("bufint_data", "bufint_coeff")


class OpSequenceTemplate:
    """Same constructor as the reference (beamform_op_sequence.py:69-83)."""

    def __init__(self, context, n_batches, n_pols, n_channels_per_stream, n_channels, n_blocks,
                 n_samples_per_block, n_ants, n_beams, xeng_id, sample_period, n_samples_per_channel) -> None:
        if n_pols != 2:
            raise ValueError("n_pols must be 2 (the reference hard-codes it: prebeamform_reorder.py:54)")
        self.context = context
        self.preBeamformReorder_template = PreBeamformReorderTemplate(
            context, n_ants, n_channels_per_stream, n_samples_per_channel, n_batches
        )
        self.beamform_coeff_template = CoeffGeneratorTemplate(
            context, n_batches, n_pols, n_channels_per_stream, n_channels, n_blocks, n_samples_per_block, n_ants,
            n_beams, xeng_id, sample_period,
        )
        self.beamform_mult_template = MatrixMultiplyTemplate(
            context=context, n_ants=n_ants, n_channels_per_stream=n_channels_per_stream,
            n_samples_per_channel=n_samples_per_channel, n_beams=n_beams, n_batches=n_batches,
        )

    def instantiate(self, queue) -> "OpSequence":
        return OpSequence(self, queue)


class OpSequence(accel.OperationSequence):
    """Fused reorder -> steering coefficients -> beamform.

    Attributes (all optional, set after ``instantiate``):
        materialize_intermediates: also fill ``bufint_data`` / ``bufint_coeff`` (default False).
        signed_input: bytes are int8 instead of the reference API's uint8 (default False).
        fp16_coeff: single fp16 rounding of the coefficients instead of the fp16 hi+lo pair (default False).
        fused: False runs the three stand-alone kernels like the reference (default True).
        batch_times: per-batch time offsets (s) for time-varying steering, see CoeffGenerator (default None).
        beam_weights: (n_beams, n_ants) real weights folded into the coefficients, see CoeffGenerator (default None).
    """

    def __init__(self, template: OpSequenceTemplate, queue) -> None:
        self.prebeamform_reorder = template.preBeamformReorder_template.instantiate(queue)
        self.beamform_coeff = template.beamform_coeff_template.instantiate(queue)
        self.beamform_mult = template.beamform_mult_template.instantiate(queue)

        operations = [
            ("prebeamform_reorder", self.prebeamform_reorder),
            ("beamform_coeff", self.beamform_coeff),
            ("beamform_mult", self.beamform_mult),
        ]
        compounds = {
            "bufin_delay_vals": ["beamform_coeff:delay_vals"],
            "bufint_coeff": ["beamform_coeff:outCoeffs", "beamform_mult:inCoeffs"],
            "bufin_reorder": ["prebeamform_reorder:inSamples"],
            "bufint_data": ["prebeamform_reorder:outReordered", "beamform_mult:inData"],
            "bufout_mult": ["beamform_mult:outData"],
        }
        super().__init__(queue, operations, compounds)
        self.template = template
        self.materialize_intermediates = False
        self.signed_input = False
        self.fp16_coeff = False
        self.fused = True
        self.batch_times = None
        self.sample_dt = 0.0
        self.beam_weights = None
        self._packed = None  # steering coefficients of the bound delay model in the tensor cores' layout (pack_coefficients)
        # The fused path never needs the reordered voltages or the coefficients in HBM, so ensure_all_bound() leaves
        # them out; code written against the reference may still ask for them (its own test reads
        # op.beamform_mult.buffer("inData").shape, beamform_op_sequence_test.py:182): they are then allocated on the
        # spot, and from the next call on they are filled as well.
        for name in _INTERMEDIATES:
            compound = self.slots[name]

            def allocate(compound=compound):
                compound.allocate(self._context())

            compound.on_demand = allocate
            for child in compound.children:
                child.on_demand = allocate

    # -- binding policy ---------------------------------------------------------------------------
    def _needs(self, name: str) -> bool:
        return name not in _INTERMEDIATES or self.materialize_intermediates or not self.fused

    def ensure_all_bound(self) -> None:
        for name in self.slots:
            if self._needs(name):
                self.ensure_bound(name)

    def check_all_bound(self) -> None:
        for name, slot in self.slots.items():
            if self._needs(name) and not slot.is_bound:
                raise ValueError(f"slot {name} is not bound")

    def flags(self) -> int:
        return (_capi.FLAG_SIGNED_INPUT if self.signed_input else 0) | (_capi.FLAG_FP16_COEFF if self.fp16_coeff else 0)

    # -- packed steering coefficients (extension) -----------------------------------------------------
    def pack_coefficients(self) -> bool:
        """Evaluate the steering coefficients of the delay model now bound to ``bufin_delay_vals`` ONCE
        (``dcbf_fused_pack_coeffs``): until ``release_coefficients()`` every call of the sequence loads them instead of
        regenerating them per heap -- bit-identical beams, a fraction of the SM-side work.  The delay model changes at
        control-plane cadence; call this again after writing a new one.  Returns False (and changes nothing) for shapes
        that keep no whole tile set (many antennas x beams) and for weights that need a scale, which stay on the
        per-call path; time-varying steering (``batch_times``) always regenerates."""
        import torch

        r = self.template.preBeamformReorder_template
        c = self.template.beamform_coeff_template
        flags = _capi.FLAG_FP16_COEFF if self.fp16_coeff else 0
        nbytes = _capi.fused_packed_bytes(r.n_ants, r.n_channels_per_stream, c.n_beams, flags)
        weights = _device_weights(self)
        if not nbytes or (weights is not None and weights_log2(self) != 0):
            self._packed = None
            return False
        dv = self.buffer("bufin_delay_vals").buffer
        with torch.cuda.stream(self.command_queue.stream):  # (the block belongs to the stream whose launches read it: a
            packed = torch.empty(nbytes, dtype=torch.uint8, device=dv.device)  # replaced pack is reused in stream order)
        _capi.fused_pack_coeffs(dv, packed, r.n_ants, r.n_channels_per_stream, c.n_channels, c.n_beams, c.xeng_id,
                                c.sample_period, flags, self.command_queue.stream, weights=weights)
        self._packed = (packed, flags)
        return True

    def release_coefficients(self) -> None:
        self._packed = None

    # -- execution --------------------------------------------------------------------------------
    def _run(self) -> None:
        self.beamform_coeff.batch_times = self.batch_times
        weights = _device_weights(self)
        self.beamform_coeff.beam_weights = weights
        if not self.fused:
            self.beamform_mult.signed_input = self.signed_input
            super()._run()
            return
        r = self.template.preBeamformReorder_template
        c = self.template.beamform_coeff_template
        if self._packed is not None and self.batch_times is None and self._packed[1] == (self.flags() & _capi.FLAG_FP16_COEFF):
            _capi.fused_packed(
                self.buffer("bufin_reorder").buffer, self._packed[0], self.buffer("bufout_mult").buffer, r.n_batches,
                r.n_ants, r.n_channels_per_stream, c.n_channels, r.n_samples_per_channel, c.n_beams, c.xeng_id,
                c.sample_period, self.flags(), self.command_queue.stream,
            )
            if self.slots["bufint_data"].is_bound:
                self.prebeamform_reorder()
            if self.slots["bufint_coeff"].is_bound:
                self.beamform_coeff()
            return
        _capi.fused(
            self.buffer("bufin_reorder").buffer, self.buffer("bufin_delay_vals").buffer,
            self.buffer("bufout_mult").buffer, r.n_batches, r.n_ants, r.n_channels_per_stream, c.n_channels,
            r.n_samples_per_channel, c.n_beams, c.xeng_id, c.sample_period, self.flags(), self.command_queue.stream,
            batch_dt=self.batch_times, weights=weights, sample_dt=self.sample_dt if self.batch_times is not None else 0.0,
            weights_log2=weights_log2(self),
        )
        if self.slots["bufint_data"].is_bound:
            self.prebeamform_reorder()
        if self.slots["bufint_coeff"].is_bound:
            self.beamform_coeff()


class QuantisedOpSequenceTemplate(OpSequenceTemplate):
    """Extension (not in the reference): the fused path with the beam post-stage -- per-beam gain, clip,
    round to int8 -- folded into the kernel's epilogue (``dcbf_fused_q8``).  Same constructor as
    ``OpSequenceTemplate``."""

    def instantiate(self, queue) -> "QuantisedOpSequence":
        return QuantisedOpSequence(self, queue)


class QuantisedOpSequence(accel.Operation):
    """.. rubric:: Slots

    bufin_reorder: (n_batches, n_ants, n_channels_per_stream, n_samples_per_channel, n_pols, 2), uint8
    bufin_delay_vals: (n_channels_per_stream, n_beams, n_ants, 4), float32
    bufin_gains: (n_beams,), float32 -- quantisation gain per beam
    bufout_q8: (n_batches, n_pols, n_channels_per_stream, n_blocks, n_samples_per_block, 2*n_beams), int8
        = clip(rint(beam * gain), -127, 127)

    Attributes: ``signed_input``, ``fp16_coeff``, ``batch_times`` as on ``OpSequence``; ``saturated`` holds the
    number of clipped values of the last call once the queue has been synchronised.
    """

    def __init__(self, template: QuantisedOpSequenceTemplate, queue) -> None:
        import numpy as np

        super().__init__(queue)
        self.template = template
        r, c, m = (template.preBeamformReorder_template, template.beamform_coeff_template,
                   template.beamform_mult_template)
        dim = accel.Dimension
        self.slots["bufin_reorder"] = accel.IOSlot(dimensions=r.inputDataShape, dtype=np.uint8)
        self.slots["bufin_delay_vals"] = accel.IOSlot(dimensions=c.delay_vals_data_dimensions, dtype=np.float32)
        self.slots["bufin_gains"] = accel.IOSlot(dimensions=(dim(c.n_beams, exact=True),), dtype=np.float32)
        self.slots["bufout_q8"] = accel.IOSlot(dimensions=m.output_data_dimensions, dtype=np.int8)
        self.signed_input = False
        self.fp16_coeff = False
        self.batch_times = None
        self._saturated = None
        self._packed = None

    def pack_coefficients(self) -> bool:
        """As ``OpSequence.pack_coefficients``: the bound delay model AND gains evaluated once
        (``dcbf_fused_pack_coeffs_q8``); pack again after writing either.  False for shapes without a whole tile set."""
        import torch

        r = self.template.preBeamformReorder_template
        c = self.template.beamform_coeff_template
        flags = _capi.FLAG_FP16_COEFF if self.fp16_coeff else 0
        nbytes = _capi.fused_packed_bytes(r.n_ants, r.n_channels_per_stream, c.n_beams, flags)
        if not nbytes:
            self._packed = None
            return False
        dv = self.buffer("bufin_delay_vals").buffer
        with torch.cuda.stream(self.command_queue.stream):  # (the block belongs to the stream whose launches read it: a
            packed = torch.empty(nbytes, dtype=torch.uint8, device=dv.device)  # replaced pack is reused in stream order)
        _capi.fused_pack_coeffs_q8(dv, self.buffer("bufin_gains").buffer, packed, r.n_ants, r.n_channels_per_stream,
                                   c.n_channels, c.n_beams, c.xeng_id, c.sample_period, flags, self.command_queue.stream)
        self._packed = (packed, flags)
        return True

    def release_coefficients(self) -> None:
        self._packed = None

    @property
    def saturated(self) -> int:
        return 0 if self._saturated is None else int(self._saturated.item())

    def _run(self) -> None:
        import torch

        r = self.template.preBeamformReorder_template
        c = self.template.beamform_coeff_template
        gains = self.buffer("bufin_gains").buffer
        if self._saturated is None:
            self._saturated = torch.zeros(1, dtype=torch.int64, device=gains.device)
        with torch.cuda.stream(self.command_queue.stream):
            self._saturated.zero_()
        flags = (_capi.FLAG_SIGNED_INPUT if self.signed_input else 0) | (_capi.FLAG_FP16_COEFF if self.fp16_coeff else 0)
        if self._packed is not None and self.batch_times is None and self._packed[1] == (flags & _capi.FLAG_FP16_COEFF):
            _capi.fused_packed_q8(
                self.buffer("bufin_reorder").buffer, self._packed[0], gains, self.buffer("bufout_q8").buffer, r.n_batches,
                r.n_ants, r.n_channels_per_stream, c.n_channels, r.n_samples_per_channel, c.n_beams, c.xeng_id,
                c.sample_period, flags, self.command_queue.stream, saturated=self._saturated,
            )
            return
        _capi.fused_q8(
            self.buffer("bufin_reorder").buffer, self.buffer("bufin_delay_vals").buffer, gains,
            self.buffer("bufout_q8").buffer, r.n_batches, r.n_ants, r.n_channels_per_stream, c.n_channels,
            r.n_samples_per_channel, c.n_beams, c.xeng_id, c.sample_period, flags, self.command_queue.stream,
            batch_dt=self.batch_times, saturated=self._saturated,
        )
