"""The parallel exact left-to-right summation (csrc/fnn_exact_sum.cuh) must equal `s += x[i]` bit for bit
on every input, including the ones it cannot collapse (ties, binade crossings, negatives, subnormals)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def seq(x):
    return np.cumsum(np.asarray(x, dtype=np.float64))[-1] if len(x) else 0.0  # numpy cumsum is strictly sequential


def cases():
    rng = np.random.default_rng(123)
    out = {}
    for n in (1, 3, 4, 5, 31, 100, 4095, 4096, 4097, 20000, 100003):
        out[f"uniform{n}"] = rng.random(n)
        out[f"lognormal{n}"] = np.exp(rng.normal(0, 3, n))
    out["ints"] = rng.integers(0, 7, 30000).astype(np.float64)
    out["halves"] = rng.integers(0, 9, 30000) * 0.5
    out["dyadic_ties"] = rng.integers(1, 2**20, 20000) * 2.0**-30 + 1.0   # many exact ties once s is large
    out["tie_storm"] = np.concatenate([[2.0**52], np.full(5000, 0.5), np.full(5000, 1.5)])
    out["tiny_after_big"] = np.concatenate([[1e300], rng.random(10000) * 1e280])
    out["zeros"] = np.zeros(9000)
    out["zeros_then"] = np.concatenate([np.zeros(5000), rng.random(5000)])
    out["negatives"] = rng.normal(0, 1, 20000)
    out["sparse_negatives"] = np.where(rng.random(20000) < 0.001, -1.0, 1.0) * rng.random(20000)
    out["subnormals"] = np.concatenate([rng.random(3000) * 1e-310, rng.random(3000)])
    out["huge_range"] = 10.0 ** rng.uniform(-200, 200, 20000)
    out["pow2"] = 2.0 ** rng.integers(-40, 40, 20000)
    out["crossing_exact"] = np.concatenate([[1.0], np.full(4096, 2.0**-52), [1.0], np.full(5000, 2.0**-52)])
    out["increasing"] = np.arange(1, 30001, dtype=np.float64) * 1.1
    out["nextafter"] = np.nextafter(np.full(20000, 1.0), 2.0)
    return out


@pytest.mark.parametrize("serial", [False, True])
def test_seq_sum_bit_exact(fnn, serial):
    for name, x in cases().items():
        ref = seq(x)
        got = fnn.seq_sum(x[None, :], serial=serial)[0]
        assert got == ref or (np.isnan(got) and np.isnan(ref)), (name, serial, got, ref)


def test_four_rows_at_once(fnn):
    rng = np.random.default_rng(5)
    rows = np.stack([rng.random(12345), np.exp(rng.normal(0, 2, 12345)), rng.integers(0, 3, 12345) * 0.25, rng.normal(0, 1, 12345)])
    got = fnn.seq_sum(rows)
    for r in range(4):
        assert got[r] == seq(rows[r]), r


def test_distance_like_rows(fnn):
    """Rows as ComputeRx sees them: distances, half of them halved."""
    from helpers import tree_matrix
    D = tree_matrix(3000, 4, 0.05)
    w = np.where(np.arange(3000) % 3 == 0, 0.5, 1.0)
    rows = np.stack([D[7] * w, D[100] * w, D[2999], D[1234] * w])
    got = fnn.seq_sum(rows)
    for r in range(4):
        assert got[r] == seq(rows[r])
