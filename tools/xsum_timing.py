"""tools/xsum_timing.py : phase cycle counts of the exact left-to-right summation (library built with -DFNN_XSUM_TIMING, FNN_LIB=...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fastneighbornet_b200 as f
rng = np.random.default_rng(1)
for m in (20000, 10000, 5000):
    x = rng.random((1, m)) + 0.3
    s = f.seq_sum(x)
    ref = 0.0
    for v in x[0]:
        ref += v
    print("m", m, "exact", s[0] == ref, flush=True)
