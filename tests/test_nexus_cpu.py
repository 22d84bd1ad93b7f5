"""CPU-only: the native Nexus writer (csrc/fnn_nexus.cpp) against a line-by-line Python restatement of
OutputPrinter.java:8-96 fed by the BitSet construction of FastNN.java:405-419 / :455-466."""
import io
import math

import numpy as np
import pytest

import fastneighbornet_b200 as fnn
from helpers import tree_matrix


def java_double(v):
    """Double.toString from Python's shortest repr (spec: JDK 19+, java.lang.Double#toString)."""
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "Infinity" if v > 0 else "-Infinity"
    sign = "-" if math.copysign(1.0, v) < 0 else ""
    v = abs(v)
    if v == 0:
        return sign + "0.0"
    digits, e10 = _shortest_digits(v)
    if -3 <= e10 < 7:
        if e10 >= 0:
            ip = (digits + "0" * (e10 + 1))[:e10 + 1]
            fp = digits[e10 + 1:] or "0"
            return sign + ip + "." + fp
        return sign + "0." + "0" * (-e10 - 1) + digits
    return sign + digits[0] + "." + (digits[1:] or "0") + "E" + str(e10)


def _shortest_digits(v):
    r = repr(v)
    if "e" in r:
        m, e = r.split("e")
        e = int(e)
    else:
        m, e = r, 0
    ip, _, fp = m.partition(".")
    if fp == "0":
        fp = ""
    digs = (ip + fp).lstrip("0")
    lead = len((ip + fp)) - len(digs)            # zeros stripped in front (0.00123 -> 2 + the integer zero)
    e10 = e + len(ip) - 1 - lead
    digs = digs.rstrip("0") or "0"
    if len(digs) == 1:                           # spec: at least two digits, the closest such decimal (4.9E-324)
        m, e = ("%.1e" % v).split("e")
        return m.replace(".", "").rstrip("0") or "0", int(e)
    return digs, e10


def reference_nexus(n, names, D, ordering, x, threshold=1e-6):
    """OutputPrinter.NexusWithSplitsAndDistances on the splitList of FastNN.java:405-419."""
    out = io.StringIO()
    P = lambda s="": out.write(s + "\n")
    split_list = []
    for i in range(n):
        cur = set()
        for j in range(i + 1, n):
            cur.add(int(ordering[j]))
            split_list.append(frozenset(cur))
    kept = [(s, x[k]) for k, s in enumerate(split_list) if x[k] > threshold]
    P("#nexus"); P()
    P("BEGIN Taxa;"); P(f"DIMENSIONS ntax={n};"); P("TAXLABELS")
    for i in range(n):
        P(f"[{i + 1}] '{names[i]}'")
    P(";"); P("END; [Taxa]"); P()
    if D is not None:
        P("BEGIN Distances;"); P(f"DIMENSIONS ntax={n};"); P("FORMAT labels=no diagonal triangle=both;"); P("MATRIX")
        for i in range(n):
            out.write("".join(" " + java_double(float(D[i, j])) for j in range(n)) + "\n")
        P(";"); P("END; [Distances]"); P()
    P("BEGIN Splits;"); P(f"DIMENSIONS ntax={n} nsplits={len(kept)};")
    P("FORMAT labels=no weights=yes confidences=no intervals=no;"); P("PROPERTIES fit=-1.0 cyclic;")
    out.write("CYCLE" + "".join(f" {int(t)}" for t in ordering[1:]) + ";\n")
    P("MATRIX")
    for c, (s, w) in enumerate(kept, 1):
        size = min(len(s), n - len(s))
        out.write(f"[{c}, size={size}] \t {java_double(float(w))} \t " + "".join(f" {t}" for t in sorted(s)) + ",\n")
    P(";"); P("END; [Splits]"); P()
    P("BEGIN st_Assumptions;"); P("uptodate;"); P("disttransform=NeighborNet;"); P("splitstransform=EqualAngle;")
    P(f"SplitsPostProcess filter=dimension value={n};"); P(" exclude  no missing;"); P("autolayoutnodelabels;")
    P("END; [st_Assumptions]"); P()
    return out.getvalue()


def test_double_to_string_documented_examples_and_random():
    # examples from the java.lang.Double#toString documentation and well-known values
    known = {1.0: "1.0", 0.0: "0.0", 100.0: "100.0", 0.001: "0.001", 1.0e-4: "1.0E-4", 1.0e7: "1.0E7", 9999999.0: "9999999.0",
             1234567.125: "1234567.125", 12345678.5: "1.23456785E7", 0.1: "0.1", 0.3: "0.3", 0.1 + 0.2: "0.30000000000000004",
             4.9e-324: "4.9E-324", 1.7976931348623157e308: "1.7976931348623157E308", 2.0 ** -44: "5.684341886080802E-14", 1.0e23: "1.0E23", 1e-323: "9.9E-324", 2e-3: "0.002", 5e-5: "5.0E-5",
             2.2250738585072014e-308: "2.2250738585072014E-308", 3.0e10: "3.0E10", 0.00999: "0.00999", 123.456: "123.456", -2.5: "-2.5",
             float("inf"): "Infinity", float("-inf"): "-Infinity"}
    for v, s in known.items():
        assert fnn.java_double_str(v) == s, (v, s)
    assert fnn.java_double_str(float("nan")) == "NaN" and fnn.java_double_str(-0.0) == "-0.0"
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.random(3000), rng.random(3000) * 10.0 ** rng.integers(-12, 13, 3000), rng.integers(0, 10 ** 8, 500).astype(float),
                           np.round(rng.random(500) * 100, 3)])
    for v in vals.tolist():
        s = fnn.java_double_str(v)
        assert s == java_double(v)
        assert float(s) == v


@pytest.mark.parametrize("n,with_D,threads", [(12, True, 1), (37, True, 3), (60, False, 0)])
def test_nexus_bytes_equal_the_restated_printer(tmp_path, n, with_D, threads):
    rng = np.random.default_rng(n)
    D = tree_matrix(n, 5, 0.05)
    ordering = np.concatenate([[0, 1], rng.permutation(n - 1) + 2]).astype(np.int32)
    npairs = n * (n - 1) // 2
    x = np.where(rng.random(npairs) < 0.15, rng.random(npairs) * 10.0 ** rng.integers(-5, 3, npairs), 0.0)
    names = [f"taxon_{i}" for i in range(n)]
    # compact (i, j, weight) triples, as fnn_weighted_splits returns them
    si, sj, w, k = [], [], [], 0
    for i in range(n):
        for j in range(i + 1, n):
            if x[k] > 1e-6:
                si.append(i); sj.append(j); w.append(x[k])
            k += 1
    p = tmp_path / "out.nex"
    fnn.write_nexus(p, ordering, si, sj, w, D=D if with_D else None, names=names, threads=threads)
    assert p.read_text() == reference_nexus(n, names, D if with_D else None, ordering, x)


def test_nexus_default_names_empty_split_list_and_errors(tmp_path):
    n = 5
    ordering = np.array([0, 1, 3, 5, 2, 4], dtype=np.int32)
    p = tmp_path / "e.nex"
    fnn.write_nexus(p, ordering, [], [], [])
    txt = p.read_text()
    assert "[3] 't3'" in txt and "nsplits=0;" in txt and "CYCLE 1 3 5 2 4;" in txt and "BEGIN Distances" not in txt
    with pytest.raises(fnn.FastNNError, match="outside"):
        fnn.write_nexus(p, ordering, [2], [2], [1.0])
    with pytest.raises(fnn.FastNNError, match="taxon id"):
        fnn.write_nexus(p, np.array([0, 1, 2, 3, 4, 9], dtype=np.int32), [], [], [])
    with pytest.raises(fnn.FastNNError, match="cannot create"):
        fnn.write_nexus(tmp_path / "no_such_dir" / "x.nex", ordering, [], [], [])
