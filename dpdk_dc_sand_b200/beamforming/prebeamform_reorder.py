"""Pre-beamform reorder operation (stand-alone stage 1).

API mirror of ``beamformer/beamforming/prebeamform_reorder.py`` (template :15-125, operation :128-186).
The CUDA-C/Mako kernel the reference JIT-compiles (``kernels/prebeamform_reorder_kernel.mako:37-92``) is
replaced by ``dcbf_reorder`` (csrc/reorder.cu): shared-memory-staged transpose, 128-bit coalesced loads and
stores, 64-bit indexing, bit-exact.
"""
import numpy as np

from .. import _capi
from ..katsdpsigproc import accel


class _KernelHandle:
    """What ``template.kernel`` holds instead of a JIT-compiled PyCUDA function."""

    def __init__(self, symbol: str) -> None:
        self.symbol = symbol

    def __repr__(self) -> str:
        return f"<libdcbf {self.symbol}>"


class PreBeamformReorderTemplate:
    """Shape algebra of the reorder; same constructor as the reference (prebeamform_reorder.py:40-47).

    Parameters
    ----------
    context: device context (``accel.Context``); unused for shape algebra, may be ``None``.
    n_ants, n_channels_per_stream, n_samples_per_channel, n_batches: as in the reference.
    """

    def __init__(self, context, n_ants: int, n_channels_per_stream: int, n_samples_per_channel: int,
                 n_batches: int) -> None:
        self.context = context
        self.n_ants = n_ants
        self.n_channels_per_stream = n_channels_per_stream
        self.n_samples_per_channel = n_samples_per_channel
        self.n_pols = 2  # reference: hard-coded (prebeamform_reorder.py:54)
        self.n_batches = n_batches
        self._sample_bitwidth = 8
        self.complexity = 2

        self.n_samples_per_block = 128 // self._sample_bitwidth  # 16 (prebeamform_reorder.py:59)
        self.n_blocks = self.n_samples_per_channel // self.n_samples_per_block
        # The reference tests `T % n_blocks` (prebeamform_reorder.py:62), which lets e.g. T=24 through and
        # divides by zero for T<16; the intent (and what the kernel needs) is a whole number of 16-sample blocks.
        if self.n_blocks == 0 or self.n_samples_per_channel % self.n_samples_per_block != 0:
            raise ValueError(f"samples_per_channel must be divisible by {self.n_samples_per_block}.")
        for name in ("n_ants", "n_channels_per_stream", "n_batches"):
            if getattr(self, name) <= 0:
                raise ValueError(f"{name} must be positive")

        self.inputDataShape = accel.exact_dimensions(   # prebeamform_reorder.py:68-75
            self.n_batches, self.n_ants, self.n_channels_per_stream, self.n_samples_per_channel, self.n_pols,
            self.complexity)
        self.outputDataShape = accel.exact_dimensions(  # :77-85
            self.n_batches, self.n_pols, self.n_channels_per_stream, self.n_blocks, self.n_samples_per_block,
            self.n_ants, self.complexity)
        self.matrix_size = self.n_ants * self.n_channels_per_stream * self.n_samples_per_channel * self.n_pols
        # Launch-shape attributes of the reference kernel, kept for API compatibility only.
        self.threads_per_block = 1024
        self.n_blocks_x = int(np.ceil(self.matrix_size / self.threads_per_block))
        self.kernel = _KernelHandle("dcbf_reorder")

    def instantiate(self, command_queue) -> "PreBeamformReorder":
        return PreBeamformReorder(self, command_queue)


class PreBeamformReorder(accel.Operation):
    """.. rubric:: Slots

    inSamples: (n_batches, n_ants, n_channels_per_stream, n_samples_per_channel, n_pols, 2), uint8
    outReordered: (n_batches, n_pols, n_channels_per_stream, n_blocks, n_samples_per_block, n_ants, 2), uint8
    """

    def __init__(self, template: PreBeamformReorderTemplate, command_queue) -> None:
        super().__init__(command_queue)
        self.template = template
        self.slots["inSamples"] = accel.IOSlot(dimensions=template.inputDataShape, dtype=np.uint8)
        self.slots["outReordered"] = accel.IOSlot(dimensions=template.outputDataShape, dtype=np.uint8)

    def _run(self) -> None:
        t = self.template
        _capi.reorder(self.buffer("inSamples").buffer, self.buffer("outReordered").buffer, t.n_batches, t.n_ants,
                      t.n_channels_per_stream, t.n_samples_per_channel, self.command_queue.stream)
