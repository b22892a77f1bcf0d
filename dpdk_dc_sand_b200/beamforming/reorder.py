"""Pre-beamform reorder on host arrays, with the reference's call signature.

Mirror of ``beamformer/beamforming/reorder.py:46-84`` (``reorder(input_data, input_data_shape,
output_data_shape)``).  The reference implements this helper with a numba CPU loop and uses it as the checker
of its GPU kernel; here there is no CPU implementation in the product package at all -- the helper stages the
host array through device memory and runs the same ``dcbf_reorder`` kernel as ``PreBeamformReorder``
(csrc/reorder.cu), and raises if no B200 is present.  (The independent CPU checker of this repo lives in
``oracle/`` and is only imported by the tests.)
"""
import numpy as np

from .. import _capi


def reorder(input_data: np.ndarray, input_data_shape: tuple, output_data_shape: tuple) -> np.ndarray:
    """(B, A, C, T, P, X) uint8 -> (B, P, C, T//S, S, A, X) uint8 through the CUDA kernel."""
    import torch

    batches, ants, chans, samples, pols, cplx = (int(v) for v in input_data_shape)
    out_shape = tuple(int(v) for v in output_data_shape)
    if pols != 2 or cplx != 2:
        raise ValueError("n_pols and complexity must both be 2")
    if out_shape != (batches, pols, chans, samples // 16, 16, ants, cplx) or samples % 16:
        raise ValueError("output_data_shape must be (B, P, C, T//16, 16, A, 2)")
    if not torch.cuda.is_available():
        raise RuntimeError("beamforming.reorder needs a CUDA device (there is no CPU implementation)")
    host = np.ascontiguousarray(np.asarray(input_data).reshape(batches, ants, chans, samples, pols, cplx))
    if host.dtype != np.uint8:
        if host.dtype.itemsize != 1:
            raise TypeError("input_data must be an 8-bit integer array")
        host = host.view(np.uint8)
    dev_in = torch.from_numpy(host).cuda()
    dev_out = torch.empty(out_shape, dtype=torch.uint8, device=dev_in.device)
    _capi.reorder(dev_in, dev_out, batches, ants, chans, samples)
    return dev_out.cpu().numpy().view(np.asarray(input_data).dtype if np.asarray(input_data).dtype.itemsize == 1 else np.uint8)
