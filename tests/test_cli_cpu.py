"""CPU-only: the thin FastNN.main-style driver parses the reference's options and fails loudly without a GPU."""
import subprocess
import sys

import pytest

from fastneighbornet_b200 import synth
from fastneighbornet_b200.__main__ import parse_args
from helpers import tree_matrix


def test_option_names_of_the_reference():
    a = parse_args(["-distFile", "x.phy", "-threads", "4", "-mode", "random_nlogn", "-mult", "7", "-order", "-additive", "-time"])
    assert (a.distFile, a.threads, a.mode, a.mult, a.order, a.additive, a.time) == ("x.phy", 4, "random_nlogn", 7, True, True, True)
    assert parse_args(["-distFile", "x.phy"]).mode == "canonical"
    with pytest.raises(SystemExit):
        parse_args(["-distFile", "x.phy", "-mode", "ORIGINAL"])
    with pytest.raises(SystemExit):
        parse_args([])


def test_no_gpu_means_an_error_not_a_cpu_run(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    p = tmp_path / "a.phy"
    synth.write_phylip(str(p), tree_matrix(12, 1, 0.05))
    r = subprocess.run([sys.executable, "-m", "fastneighbornet_b200", "-distFile", str(p), "-order"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr and r.stdout == ""
