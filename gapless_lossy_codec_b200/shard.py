"""Multi-GPU sharding of the hot path: by file, one process per GPU, no data-path collective.

The reference parallelises over frames inside one process (rayon, src/codec.rs:462, 620); files are
independent (`glc` loops over them, src/main.rs:546-583), so across GPUs the unit of sharding is the
file (BASELINE config 4: 10 000 short tracks over 8 GPUs).  Each rank runs the batched entry points
on its own context; the only inter-rank step is a host-side gather of the encoded streams
(`torch.distributed.gather_object`, CPU tensors over gloo or the default group) -- no NCCL collective
is involved in producing any byte.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

FRAME_SIZE = 2048
HOP_SIZE = 1024


def frames_for(samples_per_channel: int) -> int:
    """Frame count of Encoder::encode for one file (src/codec.rs:433-455)."""
    padded = 512 + samples_per_channel
    rem = padded % HOP_SIZE
    if rem:
        padded += HOP_SIZE - rem
    padded += 512
    return (padded - FRAME_SIZE) // HOP_SIZE + 1


def plan_by_file(work: Sequence[int], world: int) -> List[List[int]]:
    """Assign file indices to ranks: longest-processing-time greedy on `work` (frame-channels per
    file).  Deterministic (ties -> lower index, lower rank), every file appears exactly once, and the
    per-rank lists are ascending so that each rank's outputs keep the caller's relative order."""
    if world < 1:
        raise ValueError("world must be >= 1")
    order = sorted(range(len(work)), key=lambda i: (-int(work[i]), i))
    load = [0] * world
    plan: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        plan[r].append(i)
        load[r] += int(work[i])
    for p in plan:
        p.sort()
    return plan


def plan_by_modulo(n_files: int, world: int) -> List[List[int]]:
    """file index modulo GPU count."""
    return [list(range(r, n_files, world)) for r in range(world)]


def encode_sharded(files: Sequence, channels: Sequence[int], sample_rate: int, rank: int, world: int,
                   encode_batch: Optional[Callable] = None, group=None, dst: int = 0):
    """Encode `files` across `world` ranks.  Every rank passes the same `files` list (or at least the
    entries of its own shard); rank `dst` returns the encoded streams in the callers' file order, the
    other ranks return None.

    `encode_batch(list_of_pcm, list_of_channels) -> list_of_EncodedAudio` defaults to this rank's
    CUDA encoder (`Encoder(sample_rate).encode_batch` on device LOCAL_RANK); tests inject a CPU
    stand-in so that the sharding/gather logic runs under gloo without a GPU.
    """
    import numpy as np

    work = [frames_for(len(np.asarray(f).reshape(-1)) // int(c)) * int(c) for f, c in zip(files, channels)]
    mine = plan_by_file(work, world)[rank]
    if encode_batch is None:
        import os

        from .codec import Context, Encoder

        ctx = Context(int(os.environ.get("LOCAL_RANK", rank)))
        encode_batch = Encoder(sample_rate, ctx).encode_batch
    local = encode_batch([files[i] for i in mine], [channels[i] for i in mine]) if mine else []
    payload = list(zip(mine, local))
    if world == 1:
        gathered = [payload]
    else:
        import torch.distributed as dist

        gathered = [None] * world if rank == dst else None
        dist.gather_object(payload, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out = [None] * len(files)
    for part in gathered:
        for idx, enc in part:
            out[idx] = enc
    return out
