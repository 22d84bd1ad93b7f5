"""Test-side helpers for the multi-GPU selection sharding.  The tile sequence itself is NOT modelled here: it comes from the
product's TileIter (csrc/fnn_tile_iter.h) through oracle.tile_sequence(); this file only holds the merge rule of the per-rank
partial (Q, i, j) min-locs - the reference's scan-order rule: smaller Q first, then smaller (i, j)
(NetMakerOriginal.java:208-233), which the kernels implement as `better()` on the key (i << 32 | j)."""


def merge_partials(partials):
    """partials: iterable of (Q, i, j); the winner is the first strict minimum in (i, j<i) scan order."""
    best = None
    for q, i, j in partials:
        if best is None or q < best[0] or (q == best[0] and (i, j) < (best[1], best[2])):
            best = (q, i, j)
    return best
