"""CPU, world_size 2, gloo: the host-side logic of the N>1 path - tile partition of the sharded scan and the
min-loc merge rule - exercised through a real torch.distributed rendezvous (no GPU, no compute calls)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, m, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from fastneighbornet_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank evaluates Q on ITS tiles of a shared synthetic problem and posts one partial
    rng = np.random.default_rng(7)
    A = rng.integers(0, 5, size=(m, m)).astype(np.float64)   # ties on purpose
    Q = np.tril(A + A.T, -1)
    mine = None
    tiles = list(sharding.rank_tiles(m, rank, world))
    for t in tiles:
        r0, c0 = sharding.decode_tile(t)
        for i in range(r0, min(r0 + sharding.TILE_ROWS, m)):
            for j in range(c0, min(c0 + sharding.TILE_COLS, i)):
                mine = sharding.merge_partials([mine, (Q[i, j], i, j)] if mine else [(Q[i, j], i, j)])
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, tiles))
    winner = sharding.merge_partials([p for p, _ in gathered if p is not None])
    all_tiles = sorted(t for _, ts in gathered for t in ts)
    # single-process answer: first strict minimum in (i, j<i) scan order
    ref = None
    for i in range(m):
        for j in range(i):
            if ref is None or Q[i, j] < ref[0]:
                ref = (Q[i, j], i, j)
    out.put((rank, winner == ref, all_tiles == list(range(sharding.total_tiles(m)))))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_scan_partition_and_merge_gloo():
    world, m = 2, 700
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, m, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok_w and ok_t for _, ok_w, ok_t in res), res


def test_tile_decode_covers_triangle():
    sys.path.insert(0, ROOT)
    from fastneighbornet_b200 import sharding
    for m in (5, 33, 512, 513, 1500):
        seen = np.zeros((m, m), dtype=np.int32)
        for t in range(sharding.total_tiles(m)):
            r0, c0 = sharding.decode_tile(t)
            for i in range(r0, min(r0 + sharding.TILE_ROWS, m)):
                seen[i, c0:min(c0 + sharding.TILE_COLS, i)] += 1
        assert (np.tril(seen, -1) == np.tril(np.ones((m, m), dtype=np.int32), -1)).all()
