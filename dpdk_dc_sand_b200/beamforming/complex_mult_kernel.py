"""Beamformer contraction launcher.

API mirror of ``beamformer/beamforming/complex_mult_kernel.py:103-162`` (``ComplexMultKernel.complex_mult``).
The numba kernel ``run_complex_mult`` (:11-100: 2A redundant threads per output row, no operand reuse) is
replaced by ``dcbf_beamform`` (csrc/beamform.cu).  Arguments are the torch tensors behind the DeviceArrays.
"""
from .. import _capi


class ComplexMultKernel:
    """Class for beamform complex multiplication."""

    def complex_mult(self, data_matrix, coeff_matrix, out, flags: int = 0, stream=None) -> None:
        """out[b,p,c,k,s,:] = sum_j float32(data[b,p,c,k,s,j]) * coeff[b,p,c,j,:]  (data viewed as [..., 2A]).

        ``self`` is the calling MatrixMultiply operation (the reference invokes it unbound the same way,
        matrix_multiply.py:158-163); ``stream`` defaults to that operation's queue.
        """
        n_batches, n_pols, n_chans, n_blocks, n_per_block, n_ants = (int(v) for v in data_matrix.shape[:6])
        if n_pols != 2:
            raise ValueError("n_pols must be 2")
        n_beams = int(coeff_matrix.shape[4]) // 2
        if stream is None:
            queue = getattr(self, "command_queue", None)
            stream = getattr(queue, "stream", None)
        _capi.beamform(data_matrix, coeff_matrix, out, n_batches, n_chans, n_blocks * n_per_block, n_ants, n_beams,
                       flags, stream)
