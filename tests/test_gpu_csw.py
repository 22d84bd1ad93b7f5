"""Seam B2 on the GPU: the split-weight kernels (csrc/fnn_csw.cu) against the parity ladder.
L1 (oracle/csw_l1.cpp: the GPU's formulation and reduction trees restated on the CPU) must match
BIT FOR BIT, which is stronger than the 1e-9 relative bound of the north star; L0 (the reference's own
summation order) and scipy NNLS on the explicit system of FastNN.java:409-441 agree to the
algorithm's noise floor (CG stop at 1e-8 relative residual)."""
import numpy as np
import pytest

import oracle
from oracle import pyref
from fastneighbornet_b200 import synth
from helpers import random_matrix, tree_matrix

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [4, 5, 6, 9, 31, 32, 33, 64, 100, 257])
def test_matvecs_bit_exact_vs_l1_and_close_to_l0(fnn, n):
    rng = np.random.default_rng(n)
    for v in (rng.random(n * (n - 1) // 2), rng.normal(0, 1, n * (n - 1) // 2)):
        ab = fnn.csw_matvec("ab", v, n)
        atx = fnn.csw_matvec("atx", v, n)
        assert (ab == oracle.l1_ab(n, v)).all()
        assert (atx == oracle.l1_atx(n, v)).all()
        scale = np.abs(v).sum()
        assert np.abs(ab - oracle.ab(n, v)).max() <= 1e-13 * scale
        assert np.abs(atx - oracle.atx(n, v)).max() <= 1e-13 * scale
        assert (fnn.csw_matvec("unconstrained", v, n) == oracle.unconstrained_ls(n, v)).all()


def _problem(n, seed, eps):
    D = tree_matrix(n, seed, eps)
    o, _, _ = oracle.order(D)
    return D, o, synth.upper_triangle(D)


@pytest.mark.parametrize("variant", ["default", "graph"])
@pytest.mark.parametrize("n,seed,eps", [(8, 1, 0.05), (20, 1, 0.05), (40, 2, 0.05), (60, 3, 0.2), (80, 1, 0.05), (120, 4, 0.05)])
def test_split_weights_bit_exact_vs_l1(fnn, n, seed, eps, variant):
    """default = the persistent cooperative CG kernel (k_cg_persistent), graph = the launch-per-phase path: same operand
    order everywhere, so both must reproduce the L1 oracle bit for bit, iteration counts included."""
    D, o, du = _problem(n, seed, eps)
    x, st = fnn.split_weights(o, du, variant=variant)
    d_pos = oracle.setup_d(o, du)
    x1, s1 = oracle.l1_split_weights(n, d_pos)
    assert st["cg_iters"] == s1["cg_iters"] and st["outer"] == s1["outer"] and st["inner"] == s1["inner"]
    assert (x == x1).all(), np.abs(x - x1).max()
    # relative 1e-9 (north star) holds trivially; state it anyway
    assert np.abs(x - x1).max() <= 1e-9 * max(1.0, np.abs(x1).max())
    # L0 = the reference's own summation order: same algorithm, agreement to its noise floor
    x0, _ = oracle.split_weights(n, d_pos)
    assert np.abs(x - x0).max() < 2e-3


@pytest.mark.parametrize("n", [4, 5, 6, 9, 33, 64, 120, 257])
def test_literal_matvecs_bit_exact_vs_literal_oracle(fnn, n):
    """The literal-order device path (namespace lit in csrc/fnn_csw.cu: the reference's n-1 dependent diagonals, rowsum in
    index order) against the literal CPU restatement of CircularSplitWeights.java:571-731 - bit for bit."""
    rng = np.random.default_rng(100 + n)
    for v in (rng.random(n * (n - 1) // 2), rng.normal(0, 1, n * (n - 1) // 2)):
        assert (fnn.csw_matvec("ab", v, n, variant="literal") == oracle.ab(n, v)).all()
        assert (fnn.csw_matvec("atx", v, n, variant="literal") == oracle.atx(n, v)).all()


@pytest.mark.parametrize("n,seed,eps", [(8, 1, 0.05), (20, 1, 0.05), (40, 2, 0.05), (60, 3, 0.2), (80, 1, 0.05), (120, 4, 0.05)])
def test_split_weights_literal_order_bit_exact_vs_literal_oracle(fnn, n, seed, eps):
    """a CUDA path against the LITERAL oracle (L0), not against a restatement of the GPU's own
    formulation.  Every sum in the reference's order (left-to-right norm and alpha dot through the exact summation,
    wavefront mat-vecs) => identical active-set path, identical CG iteration counts, identical weights."""
    D, o, du = _problem(n, seed, eps)
    x, st = fnn.split_weights(o, du, variant="literal")
    d_pos = oracle.setup_d(o, du)
    x0, s0 = oracle.split_weights(n, d_pos)
    assert (st["cg_iters"], st["cg_calls"], st["outer"], st["inner"]) == (s0["cg_iters"], s0["cg_calls"], s0["outer"], s0["inner"])
    assert (x == x0).all(), np.abs(x - x0).max()
    assert np.abs(x - x0).max() <= 1e-9 * max(1.0, np.abs(x0).max())   # the north star's bound, met with zero difference


@pytest.mark.parametrize("n,seed,eps", [(40, 2, 0.05), (80, 1, 0.05), (120, 4, 0.05), (200, 7, 0.05)])
def test_production_formulation_vs_literal_order(fnn, n, seed, eps):
    """What the production formulation (2-D prefix sums, fixed trees) costs against the reference's order.  The active-set
    algorithm stops at a 1e-8 relative residual and a -1e-7 gradient, so two summation orders of the SAME algorithm end on
    slightly different feasible points.  Measured here (B200, round 2): the split SETS (x > 1e-6,
    FastNN.java:455-466) are identical at n = 40..120 and differ in 10 of ~710 splits at n = 200 (all with weights < 2e-3);
    the weights differ by up to 3e-4 (n = 120) / 1.8e-3 (n = 200) absolute, a few per cent on weights near 1e-2 of the largest.  The north star's 1e-9 bound therefore refers to the literal-order path
    (test_split_weights_literal_order_bit_exact_vs_literal_oracle: zero difference); this test pins the production path to
    the algorithm's noise floor."""
    D, o, du = _problem(n, seed, eps)
    xf, _ = fnn.split_weights(o, du)
    xl, _ = fnn.split_weights(o, du, variant="literal")
    kf, kl = xf > 1e-6, xl > 1e-6
    sym = np.nonzero(kf != kl)[0]
    both = kf & kl
    rel = np.abs(xf[both] - xl[both]) / xl[both]
    big = both & (xl > 1e-2 * xl.max())
    rel_big = np.abs(xf[big] - xl[big]) / xl[big]
    print(f"n={n}: kept {int(kf.sum())} / {int(kl.sum())}, symmetric difference {sym.size}, "
          f"max rel diff over kept {rel.max():.2e}, over weights > 1e-2*max {rel_big.max():.2e}, max abs {np.abs(xf - xl).max():.2e}")
    assert np.abs(xf - xl).max() < 5e-3
    assert sym.size <= max(2, int(0.02 * kl.sum()))
    assert np.maximum(xf[sym], xl[sym]).max(initial=0.0) < 5e-3   # a split in only one set carries a noise-level weight
    assert rel_big.max() < 5e-2


def test_persistent_kernel_equals_graph_path_n300(fnn):
    """A size with several CTAs per phase and a multi-level reduction tree (np = 44850 -> 44 level-1 blocks)."""
    D, o, du = _problem(300, 9, 0.05)
    xa, sa = fnn.split_weights(o, du)
    xb, sb = fnn.split_weights(o, du, variant="graph")
    assert sa["cg_iters"] == sb["cg_iters"] and (xa == xb).all()
    assert sa["kernel_launches"] < sb["kernel_launches"] / 50


def test_split_weights_match_committed_golden_fixtures(fnn):
    """tests/golden/csw_golden.json (generated by tests/golden/make_golden.py from the oracle): the production path must
    reproduce the frozen L1 weights and the literal-order path the frozen L0 weights, bit for bit, without the oracle in
    the loop."""
    import hashlib
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "csw_golden.json")) as f:
        cases = json.load(f)
    for g in cases:
        D = tree_matrix(g["n"], g["seed"], g["eps"])
        o = fnn.order(D)
        du = synth.upper_triangle(D)
        x1, s1 = fnn.split_weights(o, du)
        x0, s0 = fnn.split_weights(o, du, variant="literal")
        assert hashlib.sha256(x1.tobytes()).hexdigest() == g["l1_sha256"] and s1["cg_iters"] == g["l1_cg_iters"]
        assert hashlib.sha256(x0.tobytes()).hexdigest() == g["l0_sha256"] and s0["cg_iters"] == g["l0_cg_iters"]
        if g["l0_weights"] is not None:
            assert x0.tolist() == g["l0_weights"]


def test_additive_tree_needs_no_iterations(fnn):
    """eps = 0: the unconstrained optimum is feasible up to rounding; weights reproduce the tree."""
    D, o, du = _problem(30, 5, 0.0)
    x, st = fnn.split_weights(o, du)
    x1, _ = oracle.l1_split_weights(30, oracle.setup_d(o, du))
    assert (x == x1).all()


def test_unconstrained_only(fnn):
    D, o, du = _problem(50, 2, 0.05)
    x = fnn.split_weights(o, du, constrained=False)[0]
    assert (x == oracle.unconstrained_ls(50, oracle.setup_d(o, du))).all()


@pytest.mark.parametrize("n,seed", [(7, 1), (10, 2), (14, 3)])
def test_against_dense_nnls(fnn, n, seed):
    """The live reference path (FastNN.java:401-453) is a dense NNLS on the explicit split system; its
    minimiser is unique, so scipy's NNLS is an independent oracle for small n."""
    from scipy.optimize import nnls
    D = random_matrix(n, seed) + 1.0
    np.fill_diagonal(D, 0.0)
    o, _, _ = oracle.order(D)
    du = synth.upper_triangle(D)
    x, _ = fnn.split_weights(o, du)
    A = np.array(pyref.live_design_matrix(n, o.tolist()))
    xs, _ = nnls(A, du, maxiter=100000)
    assert np.abs(x - xs).max() < 1e-4
    assert ((x > 1e-6) == (xs > 1e-6)).all() or np.abs(x - xs).max() < 1e-5


def test_weighted_splits_emission(fnn):
    D, o, du = _problem(25, 3, 0.05)
    x, _ = fnn.split_weights(o, du)
    splits = fnn.weighted_splits(o, x)
    assert len(splits) == int((x > 1e-6).sum())
    # split (i,j) lists ordering[i+1..j]
    assert all(len(s) >= 1 and w > 1e-6 for s, w in splits)


def test_device_compacted_split_emission(fnn):
    """fnn_weighted_splits (N2): the kept splits compacted on the device equal the x > 1e-6 filter of FastNN.java:455-466."""
    D, o, du = _problem(90, 6, 0.05)
    x, _ = fnn.split_weights(o, du)
    si, sj, w = fnn.network_splits(o, du)
    n = 90
    idx = si.astype(np.int64) * (2 * n - si - 3) // 2 + sj - 1
    keep = np.nonzero(x > 1e-6)[0]
    assert (idx == keep).all() and (w == x[keep]).all()
    ref = fnn.weighted_splits(o, x)
    assert [sorted(int(t) for t in o[i + 1: j + 1]) for i, j in zip(si, sj)] == [s for s, _ in ref]


def test_network_one_call(fnn):
    """fnn_network = fnn_order followed by fnn_weighted_splits on the same matrix, distances kept on the device."""
    D = tree_matrix(70, 8, 0.05)
    o, si, sj, w = fnn.network(D)
    o_ref = fnn.order(D)
    s2 = fnn.network_splits(o_ref, synth.upper_triangle(D))
    assert (o == o_ref).all()
    assert (si == s2[0]).all() and (sj == s2[1]).all() and (w == s2[2]).all()


def test_network_recovers_circular_split_system(fnn):
    """Known answer (consistency of Neighbor-Net on circular metrics): a metric built from strictly positive weights on all
    circular splits of a cycle gives back that cycle and those weights."""
    from helpers import canon_cycle, circular_metric, split_dict
    n = 40
    cyc, w, D, du = circular_metric(n, 5)
    o, si, sj, wt = fnn.network(D, cutoff=0.0)
    assert canon_cycle(o.tolist()) == canon_cycle(cyc)
    x = np.zeros(n * (n - 1) // 2)
    rs = lambda i: i * (2 * n - i - 1) // 2
    for i, j, v in zip(si, sj, wt):
        x[rs(i) + (j - i - 1)] = v
    got, want = split_dict(n, o, x), split_dict(n, cyc, w)
    assert max(abs(got[k] - want[k]) for k in want) < 1e-8
