"""Shared input builders for the parity tests (seeded; same bytes for oracle and GPU)."""
import numpy as np

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
