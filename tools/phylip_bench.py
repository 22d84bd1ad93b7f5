"""Times the native Phylip loader (fnn_read_phylip) on a synthetic lower-triangular file.
usage: python tools/phylip_bench.py [n] [digits]   (file is written under gpurun_out/phy/, scratch)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastneighbornet_b200 as fnn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
digits = int(sys.argv[2]) if len(sys.argv) > 2 else 6
d = os.path.join(ROOT, "gpurun_out", "phy")
os.makedirs(d, exist_ok=True)
path = os.path.join(d, f"bench_{n}_{digits}.phy")
if not os.path.exists(path):
    rng = np.random.default_rng(1)
    with open(path, "w") as f:
        f.write(f"{n}\n")
        for i in range(n):
            row = rng.random(i) * 3.0
            f.write(f"t{i + 1} " + " ".join(np.char.mod(f"%.{digits}f" if digits < 17 else "%.17g", row)) + "\n")
size = os.path.getsize(path)
vals = n * (n - 1) // 2
for threads in (1, 0):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        D, names = fnn.read_phylip(path, threads=threads)
        best = min(best, time.perf_counter() - t0)
    print(f"n={n} digits={digits} file={size / 1e6:.1f} MB threads={'all' if threads == 0 else threads}: {best:.3f} s = "
          f"{size / 1e6 / best:.0f} MB/s, {vals / 1e6 / best:.1f} M values/s")
print("path", path)
