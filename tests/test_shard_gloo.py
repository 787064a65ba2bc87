"""N>1 path on CPU: world_size-2 gloo run of the by-file sharding + host-side gather.  The per-rank
worker is the oracle (no GPU here); what is under test is the plan, the ordering and the gather."""
import os
import socket

import numpy as np
import pytest

import signals
from gapless_lossy_codec_b200 import shard


def test_plan_properties():
    rng = np.random.default_rng(3)
    work = rng.integers(100, 900, 57).tolist()
    for world in (1, 2, 3, 8):
        plan = shard.plan_by_file(work, world)
        flat = sorted(i for p in plan for i in p)
        assert flat == list(range(len(work)))            # every file exactly once
        assert all(p == sorted(p) for p in plan)
        loads = [sum(work[i] for i in p) for p in plan]
        assert max(loads) - min(loads) <= max(work)      # LPT bound
        assert plan == shard.plan_by_file(work, world)   # deterministic
    assert shard.plan_by_modulo(5, 2) == [[0, 2, 4], [1, 3]]
    assert shard.plan_by_file([], 4) == [[], [], [], []]


def test_frames_for_matches_oracle():
    import oracle

    for n in (513, 1024, 1535, 1536, 1537, 44100, 88200):
        e = oracle.encode(np.zeros(n, np.float32), 1, 44100)
        assert shard.frames_for(n) == e.n_frames


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    import oracle
    from parity import assert_encoded_equal

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        files = [signals.sine(440, 44100, 1, 0.4), signals.music_like(44100, 2, 0.3), signals.square(220, 44100, 1, 0.2),
                 signals.white_noise(44100, 2, 0.15, 7), signals.sine(880, 44100, 1, 0.1)]
        chs = [1, 2, 1, 2, 1]
        calls = []

        def cpu_batch(pcms, cc):  # CPU stand-in for Encoder.encode_batch on this rank's GPU
            calls.append(len(pcms))
            return [oracle.encode(p, c, 44100, threads=2) for p, c in zip(pcms, cc)]

        out = shard.encode_sharded(files, chs, 44100, rank, world, encode_batch=cpu_batch)
        if rank == 0:
            assert len(out) == len(files) and all(o is not None for o in out)
            for i, (f, c) in enumerate(zip(files, chs)):
                assert_encoded_equal(out[i], oracle.encode(f, c, 44100, threads=2), f"file {i}")
            # gapless property across the shard boundary: sum of decoded lengths == sum of inputs
            assert sum(len(oracle.decode(o, threads=2)) for o in out) == sum(len(f) for f in files)
        else:
            assert out is None
        q.put((rank, calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo_shard_and_gather():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    got = dict(q.get(timeout=5) for _ in range(2))
    assert sum(sum(c) for c in got.values()) == 5 and all(len(c) == 1 for c in got.values())


def test_native_planner_matches_python_planner():
    """glc_plan_shards (C++, used by the glc_*_batch_sharded calls) and shard.plan_by_file make the same plan."""
    import numpy as np

    from gapless_lossy_codec_b200 import shard

    rng = np.random.default_rng(3)
    for n, world in [(1, 1), (5, 2), (40, 8), (1000, 8), (7, 16)]:
        work = [int(v) for v in rng.integers(1, 5000, n)]
        work[0] = work[-1]  # ties
        native = shard.plan_shards_native(work, world)
        plan = shard.plan_by_file(work, world)
        want = [None] * n
        for r, files in enumerate(plan):
            for i in files:
                want[i] = r
        assert native == want, (n, world)
