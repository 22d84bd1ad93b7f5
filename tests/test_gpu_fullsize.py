"""BASELINE-size runs checked through size-independent properties of the domain (the oracle needs hours at these
sizes): the ordering is a permutation starting (0, 1, ...); on an EXACT additive tree metric every clade of the tree
is an interval of the circular order; the run is deterministic; split weights of a tree metric reproduce it (A x = d)."""
import numpy as np
import pytest

from fastneighbornet_b200 import synth

pytestmark = pytest.mark.gpu


def _clade_violations(o, h, pi):
    n = len(o) - 1
    where = np.empty(n + 1, dtype=np.int64)
    where[o[1:]] = np.arange(n)
    bad = 0
    stack = [(0, len(h))]
    while stack:   # Cartesian tree of the separator heights = the generator's tree
        l, r = stack.pop()
        if r - l < 1:
            continue
        p = np.sort(where[pi[l:r + 1] + 1])
        gaps = np.diff(p)
        wrap = p[0] + n - p[-1]
        bad += int((gaps > 1).sum() + (wrap > 1) > 1)
        k = l + int(np.argmax(h[l:r]))
        stack.append((l, k))
        stack.append((k + 1, r))
    return bad


def test_canonical_n20000_additive_tree_properties(fnn):
    n = 20000
    h, a, pi, inv = synth.tree_params(n, 11)
    with fnn.Context(n) as c:
        c.synth(11, 0.0)
        o = c.order()
        st = c.stats()
        c.synth(11, 0.0)
        o2 = c.order()
    assert o[0] == 0 and o[1] == 1
    assert (np.sort(o[1:]) == np.arange(1, n + 1)).all()
    assert 19997 <= st["iterations"] <= 19999
    assert (o == o2).all(), "not deterministic"
    assert _clade_violations(o, h, pi) == 0


def test_relaxed_n20000_is_a_permutation(fnn):
    n = 20000
    with fnn.Context(n, mode="relaxed", seed=7) as c:
        c.synth(3, 0.05)
        o = c.order()
    assert o[0] == 0 and o[1] == 1 and (np.sort(o[1:]) == np.arange(1, n + 1)).all()


def test_split_weights_reproduce_tree_metric_n300(fnn):
    n = 300
    D = synth.additive_noise_matrix(n, 5, 0.0)
    o = fnn.order(D)
    du = synth.upper_triangle(D)
    x, st = fnn.split_weights(o, du)
    assert (x >= 0).all()
    # d in circular-position order (rotated permutation) and A x on the GPU
    taxa = np.concatenate([[o[n]], o[1:n]]) - 1
    Dp = D[np.ix_(taxa, taxa)]
    d_pos = Dp[np.triu_indices(n, 1)]
    ax = fnn.csw_matvec("ab", x, n)
    assert np.abs(ax - d_pos).max() < 1e-7 * d_pos.max()
    assert int((x > 1e-6).sum()) <= 2 * n - 3   # a tree has at most 2n-3 splits
