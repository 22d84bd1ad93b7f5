"""Generates tests/golden/order_golden.json from the CPU oracle (the reference ships no golden
vectors and cannot run here; see oracle/nnet_oracle.cpp header: PARITY UNPINNED).  The fixtures
freeze the oracle's output so that neither the oracle nor the CUDA path can drift silently.
Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
import oracle  # noqa: E402
from helpers import integer_matrix, tree_matrix  # noqa: E402

CASES = [("tree", 8, 1, 0.0), ("tree", 33, 2, 0.05), ("tree", 200, 1, 0.0), ("tree", 257, 3, 0.05),
         ("int", 40, 4, None), ("int", 150, 5, None), ("tree", 1100, 6, 0.05)]


def build(kind, n, seed, eps):
    return tree_matrix(n, seed, eps) if kind == "tree" else integer_matrix(n, seed)


def main():
    out = []
    for kind, n, seed, eps in CASES:
        D = build(kind, n, seed, eps)
        o, tr, _ = oracle.order(D)
        out.append({"kind": kind, "n": n, "seed": seed, "eps": eps,
                    "input_sha256": hashlib.sha256(D.tobytes()).hexdigest(),
                    "ordering": o.tolist() if n <= 300 else None,
                    "ordering_sha256": hashlib.sha256(o.astype(np.int32).tobytes()).hexdigest(),
                    "trace_sha256": hashlib.sha256(tr.tobytes()).hexdigest(),
                    "iterations": int(tr.shape[0])})
    with open(os.path.join(HERE, "order_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out), "cases")
    # split weights: both parity-ladder levels (L0 = the reference's order, L1 = the production formulation), frozen
    from fastneighbornet_b200 import synth
    csw = []
    for n, seed, eps in ((8, 1, 0.05), (14, 2, 0.05), (24, 6, 0.05), (40, 2, 0.05), (60, 3, 0.2)):
        D = tree_matrix(n, seed, eps)
        o, _, _ = oracle.order(D)
        d_pos = oracle.setup_d(o, synth.upper_triangle(D))
        x0, s0 = oracle.split_weights(n, d_pos)
        x1, s1 = oracle.l1_split_weights(n, d_pos)
        csw.append({"n": n, "seed": seed, "eps": eps, "d_pos_sha256": hashlib.sha256(d_pos.tobytes()).hexdigest(),
                    "l0_sha256": hashlib.sha256(x0.tobytes()).hexdigest(), "l0_cg_iters": s0["cg_iters"], "l0_kept": int((x0 > 1e-6).sum()),
                    "l1_sha256": hashlib.sha256(x1.tobytes()).hexdigest(), "l1_cg_iters": s1["cg_iters"], "l1_kept": int((x1 > 1e-6).sum()),
                    "l0_weights": x0.tolist() if n <= 14 else None})
    with open(os.path.join(HERE, "csw_golden.json"), "w") as f:
        json.dump(csw, f, indent=1)
    print("wrote", len(csw), "split-weight cases")


if __name__ == "__main__":
    main()
