"""Multi-GPU = independent sequence replicas, no collective on the data path. The only cross-rank step is
bench.py's max-over-ranks timing / sum of frames; it is exercised here with gloo on CPU, world_size 2."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seeds = bench.shard_sequences(8, rank, world)
    # each rank "processes" its sequences; rank 1 is slower
    frames, seconds = 100 * len(seeds), 0.5 + 0.25 * rank
    total_frames, max_seconds = bench.reduce_over_ranks(frames, seconds, device="cpu")
    q.put((rank, seeds, total_frames, max_seconds))
    dist.destroy_process_group()


def test_shard_and_reduce_world2():
    import bench
    assert bench.shard_sequences(8, 0, 1) == list(range(8))
    parts = [bench.shard_sequences(8, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == list(range(8)) and all(len(p) == 2 for p in parts)
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5, 7]
    for r in res:
        assert r[2] == 800 and abs(r[3] - 0.75) < 1e-9
