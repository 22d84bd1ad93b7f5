"""CPU: the product's resumable state machine for the Relaxed control flow (csrc/fnn_relaxed_sm.h, the code the device
kernel k_relaxed_select runs on one lane) against the literal restatement of NeighborNetLocal.findNodes: same RNG draws,
same row-permutation bookkeeping, same cache hits, same chosen pair - orderings and traces must be identical."""
import numpy as np
import pytest

import oracle
from helpers import integer_matrix, random_matrix, tree_matrix


def _both(D, **kw):
    oracle.set_relaxed_sm(False)
    a = oracle.order(D, mode="relaxed", **kw)
    oracle.set_relaxed_sm(True)
    try:
        b = oracle.order(D, mode="relaxed", **kw)
    finally:
        oracle.set_relaxed_sm(False)
    return a, b


@pytest.mark.parametrize("additive", [False, True])
def test_state_machine_equals_literal_findnodes(additive):
    for n, fb in ((9, 4), (40, 8), (150, 8), (400, 64)):
        for seed in (1, 2, 3):
            for D in (tree_matrix(n, seed, 0.0), tree_matrix(n, seed, 0.1), random_matrix(n, seed), integer_matrix(n, seed)):
                (o1, t1, _), (o2, t2, _) = _both(D, seed=40 + seed, fallback=fb, additive=additive)
                assert (o1 == o2).all() and t1.shape == t2.shape and (t1 == t2).all(), (n, seed, additive)


def test_state_machine_default_fallback():
    D = tree_matrix(1300, 5, 0.05)
    (o1, t1, _), (o2, t2, _) = _both(D, seed=12345, fallback=1024)
    assert (o1 == o2).all() and (t1 == t2).all()
