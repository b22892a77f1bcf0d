"""Frequency-channel sharding of the beamforming path across the GPUs of one box.

The path shards with NO data-path collective: every output ``(b, p, c, ...)`` depends only on channel ``c``
of the voltages and ``delay_vals[c]``.  This is the reference's own X-engine decomposition -- each engine owns
``n_channels_per_stream = n_channels // n_engines`` contiguous channels and evaluates the steering phase at the
absolute channel ``c + n_channels_per_stream * xeng_id`` (reference: beamformer/beamforming/coeff_generator.py:53,
beamformer/unit_test/coeff_generator_cpu.py:138-141; ``n_channels_per_stream = n_channels // n_ants // 4`` in the
tests, e.g. beamformer/unit_test/prebeamform_reorder_test.py:72).  Here rank ``r`` of ``world`` *is* X-engine
``xeng_id = r``.

One process per GPU (``torch.distributed``; NCCL over NVLink on the GPUs, gloo in the CPU tests).  The only
collective is the OPTIONAL gather of the beam outputs along the channel axis to one rank (``gather_beams``);
nothing in the compute path waits on another rank.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional


@dataclass(frozen=True)
class ChannelShard:
    """What rank ``rank`` of ``world`` owns: channels ``[first_channel, first_channel + n_channels_per_stream)``."""

    rank: int
    world: int
    n_channels: int             # total F-engine channels N (the `n_channels` ctor argument of the operators)
    n_channels_per_stream: int  # C: channels this rank processes

    @property
    def xeng_id(self) -> int:
        return self.rank

    @property
    def first_channel(self) -> int:
        return self.rank * self.n_channels_per_stream

    @property
    def channels(self) -> slice:
        return slice(self.first_channel, self.first_channel + self.n_channels_per_stream)


def plan(n_channels: int, world: Optional[int] = None, rank: Optional[int] = None) -> ChannelShard:
    """Shard ``n_channels`` over ``world`` ranks (defaults: WORLD_SIZE / RANK from the torchrun environment)."""
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else int(world)
    rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if n_channels <= 0 or n_channels % world:
        # the reference's `n_channels // n_engines` silently drops the remainder; refuse instead
        raise ValueError(f"n_channels ({n_channels}) must be a positive multiple of the number of ranks ({world})")
    return ChannelShard(rank, world, n_channels, n_channels // world)


def bind_to_device_numa(device_index: int) -> Optional[str]:
    """Pin the calling thread (and, by first touch, the page-locked buffers it allocates afterwards) to the CPUs
    NVML reports as local to GPU ``device_index``.  One process per GPU: without this all ranks' host staging
    buffers tend to land on one socket and the host-buffer path (``dcbf_host_plan_*``) shares one memory controller
    and one inter-socket link.  Returns a short description, or None if NVML or the cpuset does not allow it."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        cpus = sorted(os.sched_getaffinity(0))
        return f"gpu {device_index}: {len(cpus)} cpus {cpus[0]}..{cpus[-1]}"
    except Exception:  # noqa: BLE001 - affinity is an optimisation, never a requirement
        return None


def local_samples(samples, shard: ChannelShard):
    """``(B, A, N, T, P, 2)`` full-band voltages -> this rank's contiguous ``(B, A, C, T, P, 2)`` block.

    Works on numpy arrays and torch tensors.  (In production each X-engine receives only its own channels;
    this helper exists for tests and for feeding synthetic full-band data.)"""
    part = samples[:, :, shard.channels]
    return part.contiguous() if hasattr(part, "contiguous") else part.copy()


def local_delay_vals(delay_vals, shard: ChannelShard):
    """``(N, M, A, 4)`` full-band delay models -> this rank's ``(C, M, A, 4)`` block."""
    part = delay_vals[shard.channels]
    return part.contiguous() if hasattr(part, "contiguous") else part.copy()


def make_op_sequence(context, queue, shard: ChannelShard, n_batches: int, n_ants: int, n_beams: int,
                     n_samples_per_channel: int, sample_period: float):
    """The rank-local fused operation: the reference's OpSequenceTemplate with ``xeng_id = rank``."""
    from .beamforming.beamform_op_sequence import OpSequenceTemplate

    tmpl = OpSequenceTemplate(context, n_batches, 2, shard.n_channels_per_stream, shard.n_channels,
                              n_samples_per_channel // 16, 16, n_ants, n_beams, shard.xeng_id, sample_period,
                              n_samples_per_channel)
    return tmpl.instantiate(queue)


def gather_beams(local_beams, shard: ChannelShard, dst: int = 0, group=None):
    """Optional: collect every rank's ``(B, P, C, K, S, 2M)`` beams on ``dst`` as ``(B, P, N, K, S, 2M)``.

    ``local_beams`` is a torch tensor (CUDA with the nccl backend, CPU with gloo).  Returns the full-band
    tensor on ``dst`` and ``None`` elsewhere.  This is the only collective of the package and it is outside
    the timed compute path."""
    import torch
    import torch.distributed as dist

    if shard.world == 1:
        return local_beams
    if tuple(local_beams.shape)[2] != shard.n_channels_per_stream:
        raise ValueError("local_beams axis 2 must be this rank's n_channels_per_stream")
    local_beams = local_beams.contiguous()
    me = dist.get_rank(group)
    n_b, n_p, c = local_beams.shape[0], local_beams.shape[1], shard.n_channels_per_stream
    # One grouped batch of point-to-point transfers (ncclSend / ncclRecv over NVLink / NVSwitch with the nccl backend):
    # for a fixed (batch, pol) a rank's C channels are one contiguous run of the full-band tensor, so every piece is
    # received straight into its final place -- no list of parts, no second pass over the gathered bytes on `dst`.
    if me == dst:
        full = torch.empty(local_beams.shape[:2] + (shard.n_channels,) + local_beams.shape[3:], dtype=local_beams.dtype,
                           device=local_beams.device)
        full[:, :, dst * c:(dst + 1) * c].copy_(local_beams)
        ops = [dist.P2POp(dist.irecv, full[b, p, r * c:(r + 1) * c], dist.get_global_rank(group, r) if group is not None else r,
                          group=group)
               for r in range(shard.world) if r != dst for b in range(n_b) for p in range(n_p)]
    else:
        full = None
        peer = dist.get_global_rank(group, dst) if group is not None else dst
        ops = [dist.P2POp(dist.isend, local_beams[b, p], peer, group=group) for b in range(n_b) for p in range(n_p)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return full
