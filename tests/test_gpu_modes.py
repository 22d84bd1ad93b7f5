"""Relaxed and Random selection on the device against the oracle under the same java.util.Random seed
(the reference's ThreadLocalRandom is unseedable; DESIGN.md "RNG contract"): ordering and the whole
per-iteration trace must be identical."""
import numpy as np
import pytest

import oracle
from helpers import integer_matrix, random_matrix, tree_matrix

pytestmark = pytest.mark.gpu


def _run(fnn, D, mode, seed, fallback, mult=5):
    with fnn.Context(D.shape[0], record_trace=1, mode=mode, seed=seed, canonical_fallback=fallback, mult=mult) as c:
        c.load_host(D)
        o = c.order()
        return o, c.trace()


def _check(fnn, D, mode, seed=777, fallback=8, mult=5):
    o_ref, tr_ref, _ = oracle.order(D, mode=mode, seed=seed, fallback=fallback, mult=mult)
    o, tr = _run(fnn, D, mode, seed, fallback, mult)
    assert tr.shape == tr_ref.shape
    bad = np.nonzero((tr != tr_ref).any(axis=1))[0]
    assert bad.size == 0, f"{mode}: first diverging iteration {bad[0]}: gpu={tr[bad[0]]} ref={tr_ref[bad[0]]}"
    assert (o == o_ref).all()


@pytest.mark.parametrize("mode", ["relaxed", "random_n", "random_nlogn", "random_logn"])
@pytest.mark.parametrize("n", [12, 33, 100, 300])
def test_modes_small_with_low_fallback(fnn, mode, n):
    for seed in (1, 2):
        for D in (tree_matrix(n, seed, 0.0), tree_matrix(n, seed, 0.1), random_matrix(n, seed)):
            _check(fnn, D, mode, seed=100 + seed)


@pytest.mark.parametrize("mode", ["relaxed", "random_n"])
def test_modes_with_ties(fnn, mode):
    _check(fnn, integer_matrix(120, 3), mode, seed=5)


@pytest.mark.parametrize("mode", ["relaxed", "random_logn", "random_n"])
def test_modes_default_fallback_n1500(fnn, mode):
    """Default canonical_fallback = 1024 (NetMakerOriginal.java:361): the strategy runs for m > 1024, the tail is canonical."""
    _check(fnn, tree_matrix(1500, 4, 0.05), mode, seed=12345, fallback=1024)


def test_relaxed_n6000_trace_exact(fnn):
    """A size at which the Relaxed strategy runs for ~5 000 iterations (m > 1024) with rows of thousands of candidates:
    batched row scans, tie lists, the strategy-only graph and the hand-over to the canonical tail - trace-exact."""
    _check(fnn, tree_matrix(6000, 9, 0.05), "relaxed", seed=4242, fallback=1024)


def test_random_mult(fnn):
    _check(fnn, tree_matrix(200, 2, 0.05), "random_nlogn", seed=9, mult=2)


def test_reference_style_classes_modes(fnn):
    D = tree_matrix(1200, 11)
    o_ref, _, _ = oracle.order(D, mode="relaxed", seed=12345)
    assert (fnn.NeighborNetLocal(D, 1200, 1, False, None).runNeighborNet() == o_ref).all()
    o_ref, _, _ = oracle.order(D, mode="random_logn", seed=12345, mult=3)
    assert (fnn.NeighborNetRandom(D, 1200, 1, None, "LOGN", 3).runNeighborNet() == o_ref).all()


@pytest.mark.parametrize("n,fallback,eps_list", [(40, 8, (0.0, 0.05)), (120, 8, (0.0, 0.05)), (1100, 1024, (0.0,))])
def test_relaxed_additive_lookahead(fnn, n, fallback, eps_list):
    """-additive (NeighborNetLocal.java:223-255, findAgglomeratedQ :280-466) with the intended test-node loop.
    On non-additive data nearly every check fails and the reference degenerates to one look-ahead per sampled row
    (O(m) look-aheads per iteration), so the noisy cases are kept small."""
    for eps in eps_list:
        D = tree_matrix(n, 1, eps)
        o_ref, tr_ref, _ = oracle.order(D, mode="relaxed", seed=32, fallback=fallback, additive=True)
        with fnn.Context(n, record_trace=1, mode="relaxed", seed=32, canonical_fallback=fallback, additive=1) as c:
            c.load_host(D)
            o = c.order()
            tr = c.trace()
        assert tr.shape == tr_ref.shape and (tr == tr_ref).all()
        assert (o == o_ref).all()


def test_random_n3000_exercises_rejections(fnn):
    """~3e7 draws: dozens of nextInt rejections and hundreds of i/i.nbr remaps must be replayed exactly by the
    speculative parallel walk (k_random_walk)."""
    _check(fnn, tree_matrix(3000, 6, 0.05), "random_n", seed=2024, fallback=1024)


def test_relaxed_tie_heavy_matrices(fnn):
    """Every candidate of every row ties exactly (constant matrix) / blocks of duplicate taxa: the row scans return long tie
    lists (position order), the mutual-nearest test walks them, and the pools stay inside their documented capacity
    (include/fastnn.h).  Trace-exact against the oracle."""
    n = 120
    const = np.ones((n, n)) - np.eye(n)
    rng = np.random.default_rng(3)
    base = tree_matrix(30, 2, 0.05)
    idx = np.repeat(np.arange(30), 4)            # 4 copies of each of 30 taxa: zero distances inside a block
    dup = base[np.ix_(idx, idx)]
    for D in (const, dup):
        _check(fnn, D, "relaxed", seed=11)
        _check(fnn, D, "random_n", seed=11)
