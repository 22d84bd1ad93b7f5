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


if __name__ == "__main__":
    main()
