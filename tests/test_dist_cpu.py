"""CPU, world_size 2, gloo: the N>1 path's tile partition - the PRODUCT's TileIter (csrc/fnn_tile_iter.h, compiled for the
host in the oracle library) with the kernels' (rank, CTA, grid) start/stride - and the min-loc merge rule, exercised
through a real torch.distributed rendezvous (no GPU, no compute calls)."""
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
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle
    import sharding_model as sharding
    TILE_ROWS, TILE_COLS = oracle.tile_shape()
    GRID = 5   # CTAs per rank (the kernels use the SM count; any grid must partition the same way)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank evaluates Q on ITS tiles of a shared synthetic problem and posts one partial
    rng = np.random.default_rng(7)
    A = rng.integers(0, 5, size=(m, m)).astype(np.float64)   # ties on purpose
    Q = np.tril(A + A.T, -1)
    mine = None
    tiles = [t for cta in range(GRID) for t in oracle.tile_sequence(m, rank, world, cta, GRID)]
    for r0, c0 in tiles:
        for i in range(r0, min(r0 + TILE_ROWS, m)):
            for j in range(c0, min(c0 + TILE_COLS, i)):
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
    single = sorted(t for cta in range(GRID) for t in oracle.tile_sequence(m, 0, 1, cta, GRID))   # world = 1: every tile
    out.put((rank, winner == ref, all_tiles == single and len(set(all_tiles)) == len(all_tiles)))
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


def test_tile_sequence_covers_triangle_exactly_once():
    """The product's TileIter: for any (world, grid) the tiles of all ranks and CTAs cover every entry below the diagonal
    exactly once (so the sharded scan reads the algorithmic bytes and no pair is evaluated twice or skipped)."""
    sys.path.insert(0, ROOT)
    import oracle
    TR, TC = oracle.tile_shape()
    for m in (5, 33, 512, 513, 1500, 4099):
        for world, grid in ((1, 3), (2, 5), (8, 7)):
            seen = np.zeros((m, m), dtype=np.int32)
            for rank in range(world):
                for cta in range(grid):
                    for r0, c0 in oracle.tile_sequence(m, rank, world, cta, grid):
                        for i in range(r0, min(r0 + TR, m)):
                            seen[i, c0:min(c0 + TC, i)] += 1
            assert (np.tril(seen, -1) == np.tril(np.ones((m, m), dtype=np.int32), -1)).all(), (m, world, grid)
