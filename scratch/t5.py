import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import fastneighbornet_b200 as f
from fastneighbornet_b200 import synth
from helpers import tree_matrix
import oracle
for n in [int(a) for a in sys.argv[1:]] or [100, 200, 400]:
    D = tree_matrix(n, 1, 0.05)
    o = f.order(D)
    du = synth.upper_triangle(D)
    t=time.time(); x, st = f.split_weights(o, du); dt=time.time()-t
    line = f"n={n} gpu {dt:.2f}s iters={st['cg_iters']} calls={st['cg_calls']} launches={st['kernel_launches']} us/iter={1e6*dt/max(1,st['cg_iters']):.1f} kept={(x>1e-6).sum()}"
    if n <= 200:
        t=time.time(); x1, s1 = oracle.l1_split_weights(n, oracle.setup_d(o, du)); dc=time.time()-t
        line += f" | cpu L1 {dc:.2f}s equal={(x==x1).all()}"
    print(line, flush=True)
