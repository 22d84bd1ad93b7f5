"""Parity of the CUDA ordering engine (through the C ABI) against the CPU oracle: the circular
ordering and the whole per-iteration trace (m, c, Cx, Cy, x, y, kind, best) must be bit-exact."""
import numpy as np
import pytest

import oracle
from helpers import integer_matrix, random_matrix, tree_matrix

pytestmark = pytest.mark.gpu


def _run_gpu(fnn, D, **opts):
    with fnn.Context(D.shape[0], record_trace=1, **opts) as c:
        c.load_host(D)
        o = c.order()
        return o, c.trace(), c.stats()


def _check(fnn, D, **opts):
    o_ref, tr_ref, _ = oracle.order(D)
    o, tr, st = _run_gpu(fnn, D, **opts)
    assert tr.shape == tr_ref.shape, (tr.shape, tr_ref.shape)
    bad = np.nonzero((tr != tr_ref).any(axis=1))[0]
    assert bad.size == 0, f"first diverging iteration {bad[0]}: gpu={tr[bad[0]]} ref={tr_ref[bad[0]]}"
    assert (o == o_ref).all()
    return st


@pytest.mark.parametrize("n", [4, 5, 6, 7, 8, 9, 12, 17, 33, 64, 100])
@pytest.mark.parametrize("kind", ["tree0", "tree", "int", "rand"])
def test_small(fnn, n, kind):
    for seed in (1, 2, 3):
        D = {"tree0": lambda: tree_matrix(n, seed, 0.0), "tree": lambda: tree_matrix(n, seed, 0.05),
             "int": lambda: integer_matrix(n, seed), "rand": lambda: random_matrix(n, seed)}[kind]()
        _check(fnn, D)


def test_config1_n200(fnn):
    """BASELINE configs[0]: canonical -order, 200-taxon additive tree (eps=0)."""
    _check(fnn, tree_matrix(200, 1, 0.0))


@pytest.mark.parametrize("use_graph", [0, 1])
def test_n1000(fnn, use_graph):
    _check(fnn, tree_matrix(1000, 2, 0.05), use_graph=use_graph)


def test_n1500_ties(fnn):
    _check(fnn, integer_matrix(1500, 7, hi=4))


def test_n3000(fnn):
    st = _check(fnn, tree_matrix(3000, 3, 0.05))
    assert st["iterations"] >= 2997


@pytest.mark.parametrize("opts", [{}, {"no_overlap": 1}, {"force_exact_pick": 1}, {"serial_chain": 1}, {"use_graph": 0}])
def test_production_path_without_trace(fnn, opts):
    """No trace recorded = the production configuration: the 4-candidate pick is decided by the certified parallel sums
    (exact left-to-right sums only on near-ties) and u.Sx is summed on the forked branch while the next scan runs with
    the new cluster masked.  The ordering must still be the oracle's, with every A/B switch."""
    for D in (tree_matrix(2500, 8, 0.05), random_matrix(1300, 3), integer_matrix(900, 5, hi=4)):
        o_ref, _, _ = oracle.order(D, want_trace=False)
        with fnn.Context(D.shape[0], **opts) as c:
            c.load_host(D)
            o = c.order()
            st = c.stats()
        assert (o == o_ref).all()
        if not opts.get("force_exact_pick"):
            assert st["picks_certified"] > 0


def test_certified_pick_statistics(fnn):
    """Generic data: (almost) every pick is certified; integer matrices: exact ties force the exact sums."""
    with fnn.Context(3000) as c:
        c.load_host(tree_matrix(3000, 3, 0.05))
        c.order()
        st = c.stats()
    assert st["picks_certified"] >= 0.99 * (st["picks_certified"] + st["picks_exact"])
    # a constant matrix ties every candidate exactly: the certificate must refuse and the exact sums decide
    D = np.ones((300, 300)) - np.eye(300)
    o_ref, _, _ = oracle.order(D, want_trace=False)
    with fnn.Context(300) as c:
        c.load_host(D)
        o = c.order()
        st2 = c.stats()
    assert (o == o_ref).all()
    assert st2["picks_exact"] > 0 and st2["picks_certified"] == 0


def test_row_block_upload_equals_whole_upload(fnn):
    """fnn_ctx_load_host_rows + fnn_ctx_commit_load (the per-rank upload of the N>1 path) on one GPU: three row blocks in
    arbitrary order give the ordering of a whole-matrix upload."""
    D = tree_matrix(777, 6, 0.05)
    with fnn.Context(777) as c:
        c.load_host(D)
        o_ref = c.order()
        for r0, r1 in ((500, 777), (0, 123), (123, 500)):
            c.load_host_rows(D[r0:r1], r0)
        c.commit_load()
        o = c.order()
    assert (o == o_ref).all()


def test_orderings_match_committed_golden_fixtures(fnn):
    """tests/golden/order_golden.json (tests/golden/make_golden.py): orderings and per-iteration traces frozen as hashes -
    the CUDA path against the committed vectors, without the oracle in the loop."""
    import hashlib
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "order_golden.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 7
    for g in cases:
        D = tree_matrix(g["n"], g["seed"], g["eps"]) if g["kind"] == "tree" else integer_matrix(g["n"], g["seed"])
        assert hashlib.sha256(D.tobytes()).hexdigest() == g["input_sha256"]
        o, tr, _ = _run_gpu(fnn, D)
        assert hashlib.sha256(o.astype(np.int32).tobytes()).hexdigest() == g["ordering_sha256"]
        assert hashlib.sha256(np.ascontiguousarray(tr).tobytes()).hexdigest() == g["trace_sha256"] and tr.shape[0] == g["iterations"]
        if g["ordering"] is not None:
            assert o.tolist() == g["ordering"]


def test_rowsums(fnn):
    D = tree_matrix(777, 5)
    assert (fnn.rowsums(D) == oracle.rowsums(D)).all()


def test_small_n_identity(fnn):
    for n in (1, 2, 3):
        assert fnn.order(np.zeros((n, n))).tolist() == list(range(n + 1))


def test_device_synth_matches_numpy(fnn):
    n = 513
    for eps in (0.0, 0.05):
        with fnn.Context(n) as c:
            c.synth(9, eps)
            D = c.read_matrix()
        assert (D == tree_matrix(n, 9, eps)).all()


def test_reference_style_classes(fnn):
    D = tree_matrix(300, 11)
    o_ref, _, _ = oracle.order(D)
    nn = fnn.NeighborNetCanonical(D, 300, 1, None)
    assert (nn.runNeighborNet() == o_ref).all()


def test_phylip_path(fnn, tmp_path):
    from fastneighbornet_b200 import synth
    D = tree_matrix(150, 4)
    p = tmp_path / "d.phy"
    synth.write_phylip(str(p), D)
    o_ref, _, _ = oracle.order(D)
    assert (fnn.order(phylip_path=str(p), n=150) == o_ref).all()
