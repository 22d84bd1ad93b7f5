import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
os.environ['FNN_DEBUG']='1'
import numpy as np, ctypes
import fastneighbornet_b200 as f
from helpers import tree_matrix
D = tree_matrix(20000 if len(sys.argv)<2 else int(sys.argv[1]), 4, 0.05)
n = D.shape[0]
w = np.where(np.arange(n) % 3 == 0, 0.5, 1.0)
for nrows in (1, 4):
    rows = np.stack([D[7]*w, D[100]*w, D[n-1], D[1234]*w])[:nrows]
    rows = np.ascontiguousarray(rows)
    for serial in (1, 0):
        o = f.default_opts(); o.reserved[1] = serial; o.reserved[2] = 50
        out = np.zeros(nrows)
        f.api._check(f.lib().fnn_seq_sum(ctypes.byref(o), f.api._dp(rows), nrows, n, f.api._dp(out)))
