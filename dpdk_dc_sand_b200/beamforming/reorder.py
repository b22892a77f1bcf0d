"""Host-side (numpy) pre-beamform reorder with the reference's call signature.

Mirror of ``beamformer/beamforming/reorder.py:46-84`` (``reorder(input_data, input_data_shape,
output_data_shape)``), kept because the reference ships it in the ``beamforming`` package and its tests
import it as the checker.  It is NOT used by any device operation of this package.
"""
import numpy as np


def reorder(input_data: np.ndarray, input_data_shape: tuple, output_data_shape: tuple) -> np.ndarray:
    """(B, A, C, T, P, X) -> (B, P, C, T//S, S, A, X); ``S`` is taken from ``output_data_shape[4]``."""
    batches, ants, chans, samples, pols, cplx = (int(v) for v in input_data_shape)
    n_blocks, per_block = int(output_data_shape[3]), int(output_data_shape[4])
    if n_blocks * per_block != samples:
        raise ValueError("output_data_shape does not tile n_samples_per_channel")
    view = np.asarray(input_data).reshape(batches, ants, chans, n_blocks, per_block, pols, cplx)
    out = np.empty(tuple(int(v) for v in output_data_shape), dtype=input_data.dtype)
    out[...] = view.transpose(0, 5, 2, 3, 4, 1, 6)
    return out
