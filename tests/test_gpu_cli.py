"""GPU: file in, Nexus out through the three native stages, as `python -m fastneighbornet_b200`."""
import subprocess
import sys

import numpy as np
import pytest

from fastneighbornet_b200 import synth
from helpers import tree_matrix

pytestmark = pytest.mark.gpu


def test_driver_equals_the_api_calls(fnn, tmp_path):
    n = 48
    D = tree_matrix(n, 6, 0.05)
    p = tmp_path / "a.phy"
    synth.write_phylip(str(p), D)
    r = subprocess.run([sys.executable, "-m", "fastneighbornet_b200", "-distFile", str(p), "-time"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    o, si, sj, w = fnn.network(D)
    q = tmp_path / "b.nex"
    fnn.write_nexus(q, o, si, sj, w, D=D, names=[f"t{i + 1}" for i in range(n)])
    assert r.stdout == q.read_text()
    assert "splits kept" in r.stderr
    r = subprocess.run([sys.executable, "-m", "fastneighbornet_b200", "-distFile", str(p), "-order", "-mode", "RELAXED"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "[" + ", ".join(str(int(v)) for v in fnn.order(D, mode="relaxed")) + "]"
