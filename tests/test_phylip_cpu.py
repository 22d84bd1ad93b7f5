"""CPU-only: the native Phylip loader (csrc/fnn_phylip.cpp) against the conventions of
DistancesAndNames.java:43-132 / FastNN.java:270-276.  Every value must be the correctly rounded double of its token
(what Double.valueOf returns), i.e. bit-identical to Python's float()."""
import numpy as np
import pytest

import fastneighbornet_b200 as fnn
from fastneighbornet_b200 import synth
from helpers import tree_matrix


def _write(path, n, rows, header=None, eol="\n"):
    with open(path, "w", newline="") as f:
        f.write((header if header is not None else str(n)) + eol)
        for r in rows:
            f.write(r + eol)


def _lower_rows(D, fmt, sep=" ", names=None):
    n = D.shape[0]
    return [(names[i] if names else f"t{i + 1}") + "".join(sep + fmt(D[i, j]) for j in range(i)) for i in range(n)]


def _expect(rows_tokens, n):
    D = np.zeros((n, n))
    for i, toks in enumerate(rows_tokens):
        for j in range(i):
            D[i, j] = D[j, i] = float(toks[j])
    return D


@pytest.mark.parametrize("fmt", [repr, lambda v: "%.6f" % v, lambda v: "%.15g" % v, lambda v: "%.3e" % v, lambda v: "%.17g" % v])
@pytest.mark.parametrize("threads", [1, 5])
def test_values_are_correctly_rounded(tmp_path, fmt, threads):
    n = 83
    D = tree_matrix(n, 3, 0.05) * 1.2345
    rows = _lower_rows(D, lambda v: fmt(float(v)))
    p = tmp_path / "a.phy"
    _write(p, n, rows)
    got, names = fnn.read_phylip(p, threads=threads)
    want = _expect([r.split(" ")[1:] for r in rows], n)
    assert got.tobytes() == want.tobytes()
    assert names[:3] == ["t1", "t2", "t3"] and len(names) == n


def test_same_as_the_python_reader_on_synth_files(tmp_path):
    D = tree_matrix(120, 9, 0.05)
    p = tmp_path / "b.phy"
    synth.write_phylip(str(p), D)
    got, names = fnn.read_phylip(p)
    ref, ref_names = synth.read_phylip(str(p))
    assert got.tobytes() == ref.tobytes() == D.tobytes()
    assert names == list(ref_names)


def test_square_tabs_runs_of_spaces_crlf_and_header_whitespace(tmp_path):
    n = 17
    D = np.round(tree_matrix(n, 4, 0.1), 5)
    full = [f"name_{i}  " + " \t ".join("%.5f" % D[i, j] for j in range(n)) + "  " for i in range(n)]   # square, mixed separators
    p = tmp_path / "c.phy"
    _write(p, n, full, header=f"  {n}\t ", eol="\r\n")
    got, names = fnn.read_phylip(p)
    assert got.tobytes() == D.tobytes()
    assert names[5] == "name_5"
    # tab right after a space-separated name, values separated by tabs only (ss[i].split("\t"), :70-75)
    rows = [f"x{i} " + "\t".join("%.5f" % D[i, j] for j in range(i)) for i in range(n)]
    _write(p, n, rows)
    assert fnn.read_phylip(p)[0].tobytes() == D.tobytes()


def test_extra_lines_after_n_rows_are_ignored_and_no_trailing_newline(tmp_path):
    n = 9
    D = np.round(tree_matrix(n, 2, 0.1), 4)
    rows = _lower_rows(D, lambda v: "%.4f" % v)
    p = tmp_path / "d.phy"
    _write(p, n, rows + ["garbage that is never parsed", "more"])
    assert fnn.read_phylip(p)[0].tobytes() == D.tobytes()
    with open(p, "w") as f:
        f.write(str(n) + "\n" + "\n".join(rows))            # file ends without '\n'
    assert fnn.read_phylip(p)[0].tobytes() == D.tobytes()


def test_odd_tokens_go_through_strtod(tmp_path):
    toks = ["0.1", "1e-320", "1.7976931348623157e308", "123456789012345678901234567890", "0.000000000000000000000000000001",
            "9007199254740993", "4.35", "1e23", "8.41e21", "5e-324", "+3.5", "00012.500", ".5", "5.", "1E2"]
    n = len(toks) + 1
    rows = ["a"] + [f"r{i}" + "".join(" 1" for _ in range(i - 1)) + " " + toks[i - 1] for i in range(1, n)]
    p = tmp_path / "e.phy"
    _write(p, n, rows)
    got, _ = fnn.read_phylip(p)
    for i in range(1, n):
        assert got[i, i - 1] == float(toks[i - 1]) and got[i - 1, i] == float(toks[i - 1]), toks[i - 1]


def test_errors(tmp_path):
    p = tmp_path / "f.phy"
    with pytest.raises(fnn.FastNNError):
        fnn.read_phylip(tmp_path / "missing.phy")
    _write(p, 4, ["a", "b 1", "c 1", "d 1 2 3"])              # row 2 is short
    with pytest.raises(fnn.FastNNError, match="row 2"):
        fnn.read_phylip(p)
    _write(p, 4, ["a", "b 1", "c 1 2"])                       # 3 rows for 4 taxa
    with pytest.raises(fnn.FastNNError, match="3 rows"):
        fnn.read_phylip(p)
    _write(p, 4, [], header="four")
    with pytest.raises(fnn.FastNNError, match="taxon count"):
        fnn.read_phylip(p)
    _write(p, 3, ["a", "b 1", "c 1 2"])
    D = np.empty((4, 4))
    import ctypes
    rc = fnn.lib().fnn_read_phylip(str(p).encode(), 4, D.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), None, 0, 0)
    assert rc != 0                                              # header says 3, caller says 4


def test_long_tokens_including_near_midpoint_decimals(tmp_path):
    """16-19 digit tokens take the extended-precision path; decimals that sit next to a midpoint between two doubles
    are the case that path must hand to strtod."""
    from decimal import Decimal, getcontext, ROUND_FLOOR, ROUND_CEILING
    import math
    getcontext().prec = 60
    rng = np.random.default_rng(11)
    toks = []
    xs = np.concatenate([rng.random(40000) * 10.0 ** rng.integers(-8, 9, 40000), rng.integers(1, 2 ** 62, 5000).astype(np.float64)])
    for x in xs.tolist():
        k = int(rng.integers(16, 20))
        toks.append("%.*g" % (k, x))
    for x in (rng.random(20000) * 10.0 ** rng.integers(-6, 7, 20000)).tolist():
        mid = (Decimal(x) + Decimal(math.nextafter(x, math.inf))) / 2          # exact midpoint, > 19 digits
        exp = mid.adjusted()
        q = Decimal(1).scaleb(exp - 18)                                        # 19 significant digits
        toks.append(str(mid.quantize(q, rounding=ROUND_FLOOR)))                # just below the midpoint
        toks.append(str(mid.quantize(q, rounding=ROUND_CEILING)))              # just above
    toks = [t for t in toks if "inf" not in t]
    n = 2
    while n * (n - 1) // 2 < len(toks):
        n += 1
    toks += ["1"] * (n * (n - 1) // 2 - len(toks))
    rows, k = [], 0
    for i in range(n):
        rows.append(f"t{i}" + "".join(" " + t for t in toks[k:k + i]))
        k += i
    p = tmp_path / "g.phy"
    _write(p, n, rows)
    got, _ = fnn.read_phylip(p)
    want = np.array([float(t) for t in toks])
    il = np.tril_indices(n, -1)
    assert got[il].tobytes() == want.tobytes()
