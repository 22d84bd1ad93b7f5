"""Shared input builders for the parity tests (seeded; same bytes for oracle and GPU)."""
import numpy as np

from oracle import pyref

from fastneighbornet_b200 import synth


def tree_matrix(n, seed, eps=0.05):
    return synth.additive_noise_matrix(n, seed, eps)


def integer_matrix(n, seed, hi=6):
    """Small-integer distances: exact Q ties everywhere, exercises the scan-order tie-break."""
    rng = np.random.default_rng(seed)
    A = rng.integers(1, hi, size=(n, n)).astype(np.float64)
    D = np.triu(A, 1)
    return D + D.T


def random_matrix(n, seed):
    rng = np.random.default_rng(seed)
    A = rng.random((n, n))
    D = np.triu(A, 1)
    return D + D.T


def circular_metric(n, seed):
    """Strictly positive weights on ALL circular splits of a random cycle: the cycle is then the unique circular ordering
    of the metric (up to rotation / reflection) and the weights are the unique non-negative solution."""
    rng = np.random.default_rng(seed)
    cyc = [0, 1] + (rng.permutation(n - 1) + 2).tolist()
    A = np.array(pyref.live_design_matrix(n, cyc))
    w = rng.integers(1, 64, A.shape[1]).astype(np.float64) / 64.0      # dyadic: the metric is exact in fp64
    du = A @ w
    D = np.zeros((n, n))
    D[np.triu_indices(n, 1)] = du
    return cyc, w, D + D.T, du


def canon_cycle(order):
    c = list(order[1:])
    k = c.index(1)
    c = c[k:] + c[:k]
    r = [c[0]] + c[:0:-1]
    return min(c, r)


def split_dict(n, ordering, x):
    out, k, full = {}, 0, frozenset(range(1, n + 1))
    for i in range(n):
        s = set()
        for j in range(i + 1, n):
            s.add(int(ordering[j]))
            fs = frozenset(s)
            out[fs if 1 not in fs else full - fs] = x[k]
            k += 1
    return out
