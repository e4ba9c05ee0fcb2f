"""World-size-2 gloo test (CPU) of the replica data-parallel host logic (no data-path collective)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qwen3_tts_b200.dp import max_over_ranks, run_sharded, shard_indices


def test_shard_indices_partition():
    for n in (0, 1, 5, 4096):
        for world in (1, 2, 3, 8):
            parts = [shard_indices(n, world, r) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    items = [f"utterance-{i}" for i in range(7)]
    seen = []

    def fn(s):
        seen.append(s)
        return (s.upper(), rank)
    out = run_sharded(items, fn, dist)
    slow = max_over_ranks(1.0 + rank, dist)
    q.put((rank, out, seen, slow))
    dist.destroy_process_group()


def test_run_sharded_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, out, seen, slow = q.get(timeout=120)
        res[rank] = (out, seen, slow)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out0, seen0, slow0 = res[0]
    assert res[1][0] is None
    assert [o[0] for o in out0] == [f"UTTERANCE-{i}" for i in range(7)]
    assert [o[1] for o in out0] == [i % 2 for i in range(7)]          # utterance i ran on rank i mod 2
    assert seen0 == [f"utterance-{i}" for i in (0, 2, 4, 6)] and res[1][1] == [f"utterance-{i}" for i in (1, 3, 5)]
    assert slow0 == res[1][2] == 2.0                                    # max over ranks
