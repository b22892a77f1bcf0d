"""The slice of ``katsdpsigproc.accel`` (v1.2) that dc_sand's beamformer uses, over torch CUDA memory.

Calling protocol kept (reference: beamformer/unit_test/beamform_op_sequence_test.py:105-163):

    ctx = accel.create_some_context(device_filter=lambda x: x.is_cuda, interactive=False)
    queue = ctx.create_command_queue()
    op = Template(ctx, ...).instantiate(queue)
    op.ensure_all_bound()
    dev = op.buffer("slot"); host = dev.empty_like(); dev.set(queue, host); op(); dev.get(queue, host)

PyTorch is used for device allocations (``torch.empty(..., device=...)``), pinned host arrays and streams
only; every kernel is launched through libdcbf's C ABI.  Not implemented (unused by the path): mako
program building, tuning database, OpenCL, visualisation, padded (non-exact) dimensions beyond bookkeeping.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Dict, Iterable, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np


def _torch():
    import torch

    return torch


# --------------------------------------------------------------------------------------------------
# Context / queue
# --------------------------------------------------------------------------------------------------
class Device:
    """One CUDA device."""

    def __init__(self, index: int) -> None:
        torch = _torch()
        self.index = index
        props = torch.cuda.get_device_properties(index)
        self.name = props.name
        self.compute_capability = (props.major, props.minor)
        self.is_cuda = True
        self.is_gpu = True
        self.is_cpu = False
        self.is_accelerator = False
        self.simd_group_size = 32

    @property
    def platform_name(self) -> str:
        return "CUDA"

    def make_context(self) -> "Context":
        return Context(self)

    def __repr__(self) -> str:
        return f"<Device {self.index}: {self.name} sm_{self.compute_capability[0]}{self.compute_capability[1]}>"


class AbstractContext:
    """Type-annotation anchor (reference imports it from ``katsdpsigproc.abc``)."""


class AbstractCommandQueue:
    """Type-annotation anchor."""


class Context(AbstractContext):
    """Owns allocations on one device (a torch device index instead of a PyCUDA context)."""

    def __init__(self, device: Device) -> None:
        self.device = device
        self._torch_device = _torch().device("cuda", device.index)

    # context-manager protocol: the reference wraps launches in ``with self.command_queue.context:``
    def __enter__(self) -> "Context":
        self._guard = _torch().cuda.device(self.device.index)
        self._guard.__enter__()
        return self

    def __exit__(self, *exc) -> None:
        self._guard.__exit__(*exc)

    def create_command_queue(self, profile: bool = False) -> "CommandQueue":
        return CommandQueue(self, profile=profile)

    def create_tuning_command_queue(self) -> "CommandQueue":
        return CommandQueue(self, profile=True)

    def allocate_raw(self, n_bytes: int):
        return _torch().empty(int(n_bytes), dtype=_torch().uint8, device=self._torch_device)

    def allocate(self, shape, dtype, padded_shape=None, raw=None) -> "DeviceArray":
        return DeviceArray(self, shape, dtype, padded_shape, raw)

    def allocate_pinned(self, shape, dtype, padded_shape=None) -> "HostArray":
        return HostArray(shape, dtype, padded_shape, context=self)


class CommandQueue(AbstractCommandQueue):
    """A CUDA stream."""

    def __init__(self, context: Context, profile: bool = False) -> None:
        torch = _torch()
        self.context = context
        with torch.cuda.device(context.device.index):
            self.stream = torch.cuda.Stream()
        self.profile = profile

    @property
    def cuda_stream(self) -> int:
        return self.stream.cuda_stream

    def finish(self) -> None:
        self.stream.synchronize()
        # an in-kernel watchdog abort leaves partly written beams: surface it where results become visible to the host
        from .. import _capi

        _capi.fused_status_poll()

    def flush(self) -> None:
        pass

    def enqueue_zero_buffer(self, buffer) -> None:
        with _torch().cuda.stream(self.stream):
            buffer.zero_()


def all_devices() -> List[Device]:
    torch = _torch()
    if not torch.cuda.is_available():
        return []
    return [Device(i) for i in range(torch.cuda.device_count())]


def create_some_context(interactive: bool = True, device_filter: Optional[Callable[[Device], bool]] = None) -> Context:
    """First device passing ``device_filter`` (no interactive prompt is ever shown)."""
    devs = [d for d in all_devices() if device_filter is None or device_filter(d)]
    if not devs:
        raise RuntimeError("No compute devices found")
    import os

    idx = int(os.environ.get("LOCAL_RANK", "0")) % len(devs) if "LOCAL_RANK" in os.environ else 0
    return devs[idx].make_context()


# --------------------------------------------------------------------------------------------------
# Arrays
# --------------------------------------------------------------------------------------------------
class HostArray(np.ndarray):
    """numpy array over page-locked memory (so ``set``/``get`` can be asynchronous DMA)."""

    def __new__(cls, shape, dtype, padded_shape=None, context=None):
        shape = tuple(int(s) for s in shape)
        padded_shape = shape if padded_shape is None else tuple(int(s) for s in padded_shape)
        dtype = np.dtype(dtype)
        n_bytes = int(np.prod(padded_shape, dtype=np.int64)) * dtype.itemsize
        torch = _torch()
        pin = torch.cuda.is_available()
        owner = torch.empty(max(n_bytes, 1), dtype=torch.uint8, pin_memory=pin)
        base = owner.numpy()[:n_bytes].view(dtype).reshape(padded_shape)
        view = base[tuple(slice(0, s) for s in shape)]
        obj = view.view(cls)
        obj._owner = owner
        obj.padded_shape = padded_shape
        return obj

    def __array_finalize__(self, obj) -> None:
        if obj is not None:
            self._owner = getattr(obj, "_owner", None)
            self.padded_shape = getattr(obj, "padded_shape", None)

    @classmethod
    def safe(cls, obj) -> bool:
        return isinstance(obj, cls) and getattr(obj, "_owner", None) is not None and obj.flags["C_CONTIGUOUS"]


class DeviceArray:
    """A typed, shaped view of device memory.  ``buffer`` is the torch tensor (``.data_ptr()`` is the address)."""

    def __init__(self, context: Context, shape, dtype, padded_shape=None, raw=None) -> None:
        torch = _torch()
        self.context = context
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.padded_shape = self.shape if padded_shape is None else tuple(int(s) for s in padded_shape)
        if len(self.padded_shape) != len(self.shape) or any(p < s for p, s in zip(self.padded_shape, self.shape)):
            raise ValueError("padded_shape must be at least as large as shape")
        n_bytes = int(np.prod(self.padded_shape, dtype=np.int64)) * self.dtype.itemsize
        if raw is None:
            raw = context.allocate_raw(max(n_bytes, 1))
        elif raw.numel() < n_bytes:
            raise ValueError("raw allocation too small")
        self._raw = raw
        self.n_bytes = n_bytes
        tdtype = {"uint8": torch.uint8, "int8": torch.int8, "float32": torch.float32, "float64": torch.float64,
                  "int32": torch.int32, "int16": torch.int16, "float16": torch.float16,
                  "int64": torch.int64}.get(self.dtype.name)
        if tdtype is None:
            raise TypeError(f"unsupported dtype {self.dtype}")
        self._tensor = raw[:n_bytes].view(tdtype).view(self.padded_shape) if n_bytes else raw[:0].view(tdtype)

    # ---- katsdpsigproc API ----
    @property
    def buffer(self):
        return self._tensor

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def empty_like(self) -> HostArray:
        return HostArray(self.shape, self.dtype, self.padded_shape, context=self.context)

    def zeros_like(self) -> HostArray:
        h = self.empty_like()
        h.fill(0)
        return h

    def _check(self, ary: np.ndarray) -> None:
        if tuple(ary.shape) != self.shape:
            raise ValueError(f"shape mismatch: expected {self.shape}, got {tuple(ary.shape)}")
        if np.dtype(ary.dtype) != self.dtype:
            raise TypeError(f"dtype mismatch: expected {self.dtype}, got {ary.dtype}")

    def _view(self):
        t = self._tensor
        if self.padded_shape != self.shape:
            t = t[tuple(slice(0, s) for s in self.shape)]
        return t

    def set_async(self, command_queue: CommandQueue, ary: np.ndarray) -> None:
        torch = _torch()
        self._check(ary)
        src = torch.from_numpy(np.ascontiguousarray(ary) if not isinstance(ary, HostArray) else np.asarray(ary))
        with torch.cuda.stream(command_queue.stream):
            self._view().copy_(src, non_blocking=True)

    def set(self, command_queue: CommandQueue, ary: np.ndarray) -> None:
        self.set_async(command_queue, ary)
        command_queue.finish()

    def get_async(self, command_queue: CommandQueue, ary: Optional[np.ndarray] = None) -> np.ndarray:
        torch = _torch()
        if ary is None:
            ary = self.empty_like()
        self._check(ary)
        if not ary.flags["WRITEABLE"]:
            raise ValueError("destination is read-only")
        dst = torch.from_numpy(np.asarray(ary))
        with torch.cuda.stream(command_queue.stream):
            dst.copy_(self._view(), non_blocking=True)
        return ary

    def get(self, command_queue: CommandQueue, ary: Optional[np.ndarray] = None) -> np.ndarray:
        ary = self.get_async(command_queue, ary)
        command_queue.finish()
        return ary

    def zero(self, command_queue: CommandQueue) -> None:
        command_queue.enqueue_zero_buffer(self._raw)


# --------------------------------------------------------------------------------------------------
# Slots
# --------------------------------------------------------------------------------------------------
class Dimension:
    """Size of one axis plus padding requirements (``exact=True``: no padding allowed)."""

    def __init__(self, size: int, min_padded_round: Optional[int] = None, min_padded_size: Optional[int] = None,
                 alignment: Optional[int] = None, align_dtype=None, exact: bool = False) -> None:
        self.size = int(size)
        self.min_padded_round = min_padded_round or 1
        self.min_padded_size = max(self.size, min_padded_size or 0)
        self.alignment = alignment or 1
        self.exact = bool(exact)
        if self.exact and (self.min_padded_size != self.size or self.size % self.min_padded_round):
            raise ValueError("exact dimension cannot require padding")

    def required_padded_size(self) -> int:
        if self.exact:
            return self.size
        r = self.min_padded_round
        return max(self.min_padded_size, -(-self.size // r) * r)

    def valid(self, padded_size: int) -> bool:
        return padded_size == self.size if self.exact else padded_size >= self.required_padded_size()

    def link(self, other: "Dimension") -> None:
        if self.size != other.size:
            raise ValueError("linked dimensions have different sizes")
        if self.exact or other.exact:
            self.exact = other.exact = True


def exact_dimensions(*sizes: int) -> Tuple[Dimension, ...]:
    """Unpadded dimensions of the given sizes (every slot of the beamformer operators is ``exact=True``)."""
    return tuple(Dimension(int(n), exact=True) for n in sizes)


class IOSlotBase:
    def __init__(self) -> None:
        self.buffer: Optional[DeviceArray] = None
        self.is_bound = False
        # optional: called by Operation.buffer() when the slot is asked for while unbound (the fused op-sequence leaves
        # its intermediates unallocated until somebody wants to look at them)
        self.on_demand = None

    def bind(self, buffer: Optional[DeviceArray]) -> None:
        self.buffer = buffer
        self.is_bound = buffer is not None


class IOSlot(IOSlotBase):
    """A typed array argument of an operation."""

    def __init__(self, dimensions: Sequence[Union[int, Dimension]], dtype) -> None:
        super().__init__()
        self.dimensions = tuple(d if isinstance(d, Dimension) else Dimension(d) for d in dimensions)
        self.shape = tuple(d.size for d in self.dimensions)
        self.dtype = np.dtype(dtype)

    def required_padded_shape(self) -> Tuple[int, ...]:
        return tuple(d.required_padded_size() for d in self.dimensions)

    def required_bytes(self) -> int:
        return int(np.prod(self.required_padded_shape(), dtype=np.int64)) * self.dtype.itemsize

    def validate(self, buffer: DeviceArray) -> None:
        if buffer.shape != self.shape:
            raise ValueError(f"buffer has shape {buffer.shape}, slot needs {self.shape}")
        if buffer.dtype != self.dtype:
            raise TypeError(f"buffer has dtype {buffer.dtype}, slot needs {self.dtype}")
        for d, p in zip(self.dimensions, buffer.padded_shape):
            if not d.valid(p):
                raise ValueError("buffer padding does not satisfy the slot")

    def bind(self, buffer: Optional[DeviceArray]) -> None:
        if buffer is not None:
            self.validate(buffer)
        super().bind(buffer)

    def allocate(self, context: Context, bind: bool = True) -> DeviceArray:
        buf = DeviceArray(context, self.shape, self.dtype, self.required_padded_shape())
        if bind:
            self.bind(buf)
        return buf

    def allocate_host(self, context: Context) -> HostArray:
        return HostArray(self.shape, self.dtype, self.required_padded_shape(), context=context)


class CompoundIOSlot(IOSlotBase):
    """Several slots (of identical shape and dtype) that must share one buffer."""

    def __init__(self, children: Iterable[IOSlotBase]) -> None:
        super().__init__()
        self.children = list(children)
        if not self.children:
            raise ValueError("empty compound slot")
        first = self._leaf(self.children[0])
        self.shape, self.dtype = first.shape, first.dtype
        self.dimensions = first.dimensions
        for ch in self.children[1:]:
            leaf = self._leaf(ch)
            if leaf.shape != self.shape:
                raise ValueError(f"compound slot children disagree on shape: {leaf.shape} vs {self.shape}")
            if leaf.dtype != self.dtype:
                raise TypeError("compound slot children disagree on dtype")
            for a, b in zip(self.dimensions, leaf.dimensions):
                a.link(b)

    @staticmethod
    def _leaf(slot):
        while isinstance(slot, CompoundIOSlot):
            slot = slot.children[0]
        return slot

    def required_padded_shape(self):
        return tuple(max(self._leaf(c).dimensions[i].required_padded_size() for c in self.children)
                     for i in range(len(self.shape)))

    def required_bytes(self) -> int:
        return int(np.prod(self.required_padded_shape(), dtype=np.int64)) * self.dtype.itemsize

    def bind(self, buffer: Optional[DeviceArray]) -> None:
        for ch in self.children:
            ch.bind(buffer)
        super().bind(buffer)

    def allocate(self, context: Context, bind: bool = True) -> DeviceArray:
        buf = DeviceArray(context, self.shape, self.dtype, self.required_padded_shape())
        if bind:
            self.bind(buf)
        return buf

    def allocate_host(self, context: Context) -> HostArray:
        return HostArray(self.shape, self.dtype, self.required_padded_shape(), context=context)


# --------------------------------------------------------------------------------------------------
# Operations
# --------------------------------------------------------------------------------------------------
class Operation:
    """Device operation with named buffer slots; subclasses implement ``_run``."""

    def __init__(self, command_queue: CommandQueue, allocator=None) -> None:
        self.slots: Dict[str, IOSlotBase] = OrderedDict()
        self.hidden_slots: Dict[str, IOSlotBase] = OrderedDict()
        self.command_queue = command_queue
        self.allocator = allocator

    def bind(self, **kwargs: Optional[DeviceArray]) -> None:
        for name, buffer in kwargs.items():
            if name not in self.slots:
                raise KeyError(f"no slot named {name}")
            self.slots[name].bind(buffer)

    def buffer(self, name: str) -> DeviceArray:
        slot = self.slots.get(name) or self.hidden_slots[name]
        if slot.buffer is None and slot.on_demand is not None:
            slot.on_demand()
        if slot.buffer is None:
            raise ValueError(f"slot {name} is not bound")
        return slot.buffer

    def _context(self) -> Context:
        ctx = getattr(self.command_queue, "context", None)
        if ctx is None:
            raise RuntimeError("operation was instantiated without a device command queue")
        return ctx

    def ensure_bound(self, name: str) -> None:
        slot = self.slots[name]
        if not slot.is_bound:
            slot.allocate(self._context())

    def ensure_all_bound(self) -> None:
        for name in self.slots:
            self.ensure_bound(name)
        for slot in self.hidden_slots.values():
            if not slot.is_bound:
                slot.allocate(self._context())

    def check_all_bound(self) -> None:
        for name, slot in list(self.slots.items()) + list(self.hidden_slots.items()):
            if not slot.is_bound:
                raise ValueError(f"slot {name} is not bound")

    def required_bytes(self) -> int:
        return sum(s.required_bytes() for s in list(self.slots.values()) + list(self.hidden_slots.values()))

    def parameters(self) -> Mapping[str, object]:
        return {}

    def _run(self) -> None:
        raise NotImplementedError

    def __call__(self, **kwargs: Optional[DeviceArray]) -> None:
        self.bind(**kwargs)
        self.check_all_bound()
        self._run()


class OperationSequence(Operation):
    """Runs child operations in order.  Child slots appear as ``"opname:slot"`` unless listed in
    ``compounds`` (``{new_name: ["op:slot", ...]}``), in which case they share one buffer."""

    def __init__(self, command_queue: CommandQueue, operations: Sequence[Tuple[str, Operation]],
                 compounds: Optional[Mapping[str, Sequence[str]]] = None, allocator=None) -> None:
        super().__init__(command_queue, allocator)
        self.operations: "OrderedDict[str, Operation]" = OrderedDict(operations)
        if len(self.operations) != len(operations):
            raise ValueError("operation names are not unique")
        for name, op in self.operations.items():
            for slot_name, slot in op.slots.items():
                self.slots[f"{name}:{slot_name}"] = slot
            for slot_name, slot in op.hidden_slots.items():
                self.hidden_slots[f"{name}:{slot_name}"] = slot
        for new_name, members in (compounds or {}).items():
            children = []
            for m in members:
                if m in self.slots:
                    children.append(self.slots.pop(m))
                # katsdpsigproc silently ignores names that do not exist
            if children:
                self.slots[new_name] = CompoundIOSlot(children)

    def _run(self) -> None:
        for op in self.operations.values():
            op()

    def __call__(self, **kwargs: Optional[DeviceArray]) -> None:
        self.bind(**kwargs)
        self.check_all_bound()
        self._run()
