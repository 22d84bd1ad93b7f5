import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # the built library is git-ignored: (re)build it when it is missing or older than its sources (nvcc cross-compiles
    # sm_100a without a GPU); the tests themselves never fall back to anything else
    import glob
    import shutil
    import subprocess
    so = os.path.join(ROOT, "fastneighbornet_b200", "libfastnn.so")
    srcs = glob.glob(os.path.join(ROOT, "fastneighbornet_b200", "csrc", "*")) + [os.path.join(ROOT, "include", "fastnn.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs)
    if stale and shutil.which("nvcc"):
        subprocess.check_call(["bash", os.path.join(ROOT, "build.sh")], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def fnn():
    import fastneighbornet_b200 as f
    return f
