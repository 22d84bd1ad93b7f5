"""Host-side description of the multi-GPU selection sharding (mirrors csrc/fnn_scan_tma.cuh).

The scan walks the lower triangle in tiles of TILE_ROWS x TILE_COLS; tiles are numbered band by band
(band g = 512 rows, KPB row tiles per band, g+1 column tiles each).  Rank r of `world` owns tiles
t = r, r + world, r + 2*world, ... and the per-rank partial (Q, i, j) min-locs are merged with the
reference's scan-order rule: smaller Q first, then smaller (i, j) (NetMakerOriginal.java:208-233).
Used by the CPU (gloo) tests of the N>1 path and by documentation; the kernels carry their own copy.
"""
TILE_ROWS = 32
TILE_COLS = 512
KPB = TILE_COLS // TILE_ROWS


def total_tiles(m):
    n_row_tiles = (m + TILE_ROWS - 1) // TILE_ROWS
    g_full, r_rem = divmod(n_row_tiles, KPB)
    return KPB * g_full * (g_full + 1) // 2 + r_rem * (g_full + 1)


def decode_tile(t):
    """tile index -> (first row, first column)."""
    g = 0
    while KPB * (g + 1) * (g + 2) // 2 <= t:
        g += 1
    rem = t - KPB * g * (g + 1) // 2
    return (g * KPB + rem // (g + 1)) * TILE_ROWS, (rem % (g + 1)) * TILE_COLS


def rank_tiles(m, rank, world):
    return range(rank, total_tiles(m), world)


def merge_partials(partials):
    """partials: iterable of (Q, i, j); the winner is the first strict minimum in (i, j<i) scan order."""
    best = None
    for q, i, j in partials:
        if best is None or q < best[0] or (q == best[0] and (i, j) < (best[1], best[2])):
            best = (q, i, j)
    return best
