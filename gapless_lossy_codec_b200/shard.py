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


# ------------------------------------------------------------------ one process, several devices


def encode_batch_devices(encoders: Sequence, files: Sequence, channels: Sequence[int]):
    """glc_encode_batch_sharded: the files of one call split over `encoders` (one Encoder per context,
    normally one context per GPU) by longest-processing-time-first in the library, one host thread per
    context, outputs in input order.  Returns (list of EncodedAudio, shard index per file)."""
    import ctypes as C

    import numpy as np

    from . import _ffi
    from .codec import EncodedAudio

    lib = encoders[0]._lib
    n, k = len(files), len(encoders)
    arrs = [np.ascontiguousarray(f, dtype=np.float32).reshape(-1) for f in files]
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    ns = (C.c_uint64 * n)(*[a.size for a in arrs])
    chs = (C.c_uint16 * n)(*[int(c) for c in channels])
    encs = (C.c_void_p * k)(*[e.handle.value for e in encoders])
    outs = (C.POINTER(_ffi.Encoded) * n)()
    shard_of = (C.c_uint32 * n)()
    _ffi.check(lib.glc_encode_batch_sharded(encs, k, n, ptrs, ns, chs, outs, shard_of))
    res = []
    try:
        for i in range(n):
            res.append(EncodedAudio._from_struct(outs[i].contents))
    finally:
        for i in range(n):
            if outs[i]:
                lib.glc_encoded_free(encoders[shard_of[i]].ctx.handle, outs[i])
    return res, list(shard_of)


def decode_batch_devices(decoders: Sequence, encoded: Sequence):
    """glc_decode_batch_sharded: see encode_batch_devices.  Returns (list of PCM arrays, shard per file)."""
    import ctypes as C

    import numpy as np

    from . import _ffi

    lib = decoders[0]._lib
    n, k = len(encoded), len(decoders)
    structs = [e._as_struct() for e in encoded]
    ptrs = (C.POINTER(_ffi.Encoded) * n)(*[C.pointer(s) for s in structs])
    decs = (C.c_void_p * k)(*[d.handle.value for d in decoders])
    outs = (C.POINTER(C.c_float) * n)()
    ns = (C.c_uint64 * n)()
    shard_of = (C.c_uint32 * n)()
    _ffi.check(lib.glc_decode_batch_sharded(decs, k, n, ptrs, outs, ns, shard_of))
    res = []
    for i in range(n):
        try:
            res.append(np.ctypeslib.as_array(outs[i], shape=(ns[i],)).copy() if ns[i] else np.zeros(0, np.float32))
        finally:
            lib.glc_free(decoders[shard_of[i]].ctx.handle, outs[i])
    return res, list(shard_of)


def flac_encode_batch_devices(contexts: Sequence, files: Sequence, sample_rates: Sequence[int], channels: Sequence[int],
                              level: int = 5):
    """glc_flac_encode_batch_sharded.  Returns (list of FLAC byte strings, shard per file)."""
    import ctypes as C

    import numpy as np

    from . import _ffi

    lib = contexts[0]._lib
    n, k = len(files), len(contexts)
    arrs = [np.ascontiguousarray(f, dtype=np.float32).reshape(-1) for f in files]
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    ns = (C.c_uint64 * n)(*[a.size for a in arrs])
    srs = (C.c_uint32 * n)(*[int(v) for v in sample_rates])
    chs = (C.c_uint16 * n)(*[int(c) for c in channels])
    ctxs = (C.c_void_p * k)(*[c.handle.value for c in contexts])
    outs = (C.POINTER(C.c_uint8) * n)()
    lens = (C.c_uint64 * n)()
    shard_of = (C.c_uint32 * n)()
    _ffi.check(lib.glc_flac_encode_batch_sharded(ctxs, k, n, ptrs, ns, srs, chs, int(level), outs, lens, shard_of))
    res = []
    for i in range(n):
        try:
            res.append(C.string_at(outs[i], lens[i]))
        finally:
            lib.glc_free(contexts[shard_of[i]].handle, outs[i])
    return res, list(shard_of)


def plan_shards_native(work: Sequence[int], world: int) -> List[int]:
    """glc_plan_shards (the C++ planner the *_sharded calls use): shard index per file."""
    import ctypes as C

    from . import _ffi

    lib = _ffi.load()
    n = len(work)
    w = (C.c_uint64 * max(n, 1))(*[int(v) for v in work])
    out = (C.c_uint32 * max(n, 1))()
    _ffi.check(lib.glc_plan_shards(n, w, int(world), out))
    return list(out)[:n]

