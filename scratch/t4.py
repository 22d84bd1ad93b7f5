import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, ctypes
import fastneighbornet_b200 as f
from helpers import tree_matrix
n=20000
D = tree_matrix(n, 4, 0.05)
w = np.where(np.arange(n) % 3 == 0, 0.5, 1.0)
rows = np.ascontiguousarray(np.stack([D[7]*w, D[100]*w, D[n-1], D[1234]*w]))
print(f.seq_sum(rows))
print(f.seq_sum(rows[:1]))
