"""Control-plane hook: double-buffered delay models and beam weights for a running beamformer operation.

The reference's control plane only forwards ``?beam-weights`` / delay updates to the engines
(ngkcs/ngkcs/corr3_servlet.py:140-153) and leaves open how an engine applies them.  Here an update never stalls
the data path: the new ``delay_vals`` (and optionally the per-(beam, antenna) weights) are uploaded on a side
stream into the buffer the kernels are NOT reading, and ``activate()`` switches the operation over at the next
heap boundary (the compute stream waits on the upload's event; a launch already queued keeps the old model).
PyTorch is used only for the device buffers, the pinned staging arrays and the streams.
"""
from __future__ import annotations

import numpy as np


class DelayModelUpdater:
    """Attach to an ``OpSequence`` / ``QuantisedOpSequence`` (any operation with a ``bufin_delay_vals`` slot)."""

    def __init__(self, op) -> None:
        import torch

        from .katsdpsigproc import accel

        self._torch = torch
        self.op = op
        slot = op.slots["bufin_delay_vals"]
        ctx = op.command_queue.context
        if not slot.is_bound:
            op.ensure_bound("bufin_delay_vals")
        self._buffers = [slot.buffer, accel.DeviceArray(ctx, slot.shape, slot.dtype)]
        self._active = 0
        self._pending = None  # (buffer index, weights tensor or None, event)
        self._copy_stream = torch.cuda.Stream(device=self._buffers[0].buffer.device)
        self._staging = [b.empty_like() for b in self._buffers]  # pinned host arrays, one per device buffer
        self._staging_done = [None, None]  # event after the last DMA out of each staging array
        self._weights = [None, None]
        self._retired = None  # event on the compute stream after the last launch that may read the inactive buffer

    def update(self, delay_vals: np.ndarray, beam_weights: np.ndarray = None) -> None:
        """Start uploading a new model; returns immediately.  The operation keeps using the current one."""
        torch = self._torch
        idx = 1 - self._active
        buf, host = self._buffers[idx], self._staging[idx]
        if tuple(delay_vals.shape) != buf.shape:
            raise ValueError(f"delay_vals must have shape {buf.shape}")
        if self._pending is not None:
            self._pending[2].synchronize()  # an earlier, never activated upload into the same buffer
        if self._staging_done[idx] is not None:
            self._staging_done[idx].synchronize()  # the previous DMA out of this staging array has finished
        host[...] = delay_vals
        w_dev = None
        # Kernels queued before the last activate() may still be reading this buffer (it was the active one until then):
        # the copy waits, on the device, for the point of the compute stream at which it was retired.
        if self._retired is None:
            self._retired = torch.cuda.Event()
            self._retired.record(self.op.command_queue.stream)
        self._copy_stream.wait_event(self._retired)
        with torch.cuda.stream(self._copy_stream):
            buf.buffer.copy_(torch.from_numpy(np.asarray(host)), non_blocking=True)
            if beam_weights is not None:
                w_host = torch.from_numpy(np.ascontiguousarray(beam_weights, dtype=np.float32)).pin_memory()
                w_dev = w_host.to(buf.buffer.device, non_blocking=True)
                self._keep = w_host
            event = torch.cuda.Event()
            event.record(self._copy_stream)
        self._staging_done[idx] = event
        self._pending = (idx, w_dev, event)

    def activate(self) -> bool:
        """Make the uploaded model current for every launch issued from now on.  False if nothing is pending."""
        if self._pending is None:
            return False
        idx, w_dev, event = self._pending
        stream = self.op.command_queue.stream
        # every launch queued so far reads the buffer that now becomes the inactive one: remember where they end
        self._retired = self._torch.cuda.Event()
        self._retired.record(stream)
        stream.wait_event(event)  # device-side wait: the host does not block
        self.op.bind(bufin_delay_vals=self._buffers[idx])
        if w_dev is not None:
            self.op.beam_weights = w_dev
        self._active, self._pending = idx, None
        # an operation that runs on packed steering coefficients gets the new model packed here, on the compute stream
        # behind the upload: launches already queued keep the old tile sets, later ones read the new ones
        if getattr(self.op, "_packed", None) is not None:
            self.op.pack_coefficients()
        return True
