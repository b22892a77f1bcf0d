"""Beamform multiplication operation (stand-alone stage 3).

API mirror of ``beamformer/beamforming/matrix_multiply.py`` (template :16-115, operation :118-163).
"""
import numpy as np

from ..katsdpsigproc import accel
from ..katsdpsigproc.accel import IOSlot, Operation
from .complex_mult_kernel import ComplexMultKernel


class MatrixMultiplyTemplate:
    """Same constructor as the reference (matrix_multiply.py:55-63)."""

    def __init__(self, context, n_ants: int, n_channels_per_stream: int, n_samples_per_channel: int, n_beams: int,
                 n_batches: int) -> None:
        self.context = context
        self.n_ants = n_ants
        self.n_channels_per_stream = n_channels_per_stream
        self.n_samples_per_channel = n_samples_per_channel
        self.n_batches = n_batches
        self._sample_bitwidth = 8
        self.n_pols = 2
        self.complexity = 2
        self.beams = n_beams

        self.n_samples_per_block = 128 // self._sample_bitwidth
        self.n_blocks = self.n_samples_per_channel // self.n_samples_per_block
        if self.n_blocks == 0 or self.n_samples_per_channel % self.n_samples_per_block != 0:
            raise ValueError(f"samples_per_channel must be divisible by {self.n_samples_per_block}.")
        self.length = self.n_batches * self.n_pols * self.n_channels_per_stream * self.n_blocks * self.n_samples_per_block

        b, p, c, k, spb, a, m2 = (self.n_batches, self.n_pols, self.n_channels_per_stream, self.n_blocks,
                                   self.n_samples_per_block, self.n_ants, self.beams * self.complexity)
        self.input_data_dimensions = accel.exact_dimensions(b, p, c, k, spb, a, self.complexity)  # matrix_multiply.py:86-94
        self.output_data_dimensions = accel.exact_dimensions(b, p, c, k, spb, m2)                  # :95-102
        self.coeff_data_dimensions = accel.exact_dimensions(b, p, c, 2 * a, m2)                    # :104-110

    def instantiate(self, command_queue) -> "MatrixMultiply":
        return MatrixMultiply(self, command_queue)


class MatrixMultiply(Operation):
    """.. rubric:: Slots

    inData: (n_batches, n_pols, n_channels_per_stream, n_blocks, n_samples_per_block, n_ants, 2), uint8
    inCoeffs: (n_batches, n_pols, n_channels_per_stream, 2*n_ants, 2*n_beams), float32
    outData: (n_batches, n_pols, n_channels_per_stream, n_blocks, n_samples_per_block, 2*n_beams), float32

    ``signed_input`` (attribute, default False like the reference's uint8 API) reinterprets bytes as int8.
    """

    def __init__(self, template: MatrixMultiplyTemplate, command_queue) -> None:
        super().__init__(command_queue)
        self.template = template
        self.signed_input = False
        self.slots["inData"] = IOSlot(dimensions=template.input_data_dimensions, dtype=np.uint8)
        self.slots["outData"] = IOSlot(dimensions=template.output_data_dimensions, dtype=np.float32)
        self.slots["inCoeffs"] = IOSlot(dimensions=template.coeff_data_dimensions, dtype=np.float32)

    def _run(self) -> None:
        from .. import _capi

        ComplexMultKernel.complex_mult(
            self,
            self.buffer("inData").buffer,
            self.buffer("inCoeffs").buffer,
            self.buffer("outData").buffer,
            flags=_capi.FLAG_SIGNED_INPUT if self.signed_input else 0,
        )
