"""torchrun --nproc-per-node N tests/mg_check.py : sharded-scan ordering must equal the single-GPU / oracle result."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch, torch.distributed as dist
import fastneighbornet_b200 as fnn
from helpers import tree_matrix, integer_matrix
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok = True
quick = "--quick" in sys.argv
cases = [("tree1500", tree_matrix(1500, 3, 0.05)), ("int700", integer_matrix(700, 2)),
         ("tree6000", tree_matrix(6000, 4, 0.05))]   # > 4096 active nodes: the scan is really sharded there
if not quick:
    cases.append(("tree4000", tree_matrix(4000, 5, 0.05)))
for name, D in cases:
    n = D.shape[0]
    with fnn.Context(n, device=lr, record_trace=1) as c:
        c.connect_torch()
        c.load_host(D)
        o = c.order(); tr = c.trace()
        # second run on the same wired context (mailbox tags must not collide), uploaded 1/world per rank + NVLink broadcast
        c.load_host_sharded(D)
        o2 = c.order()
    with fnn.Context(n, device=lr, record_trace=1) as c1:   # un-wired single-GPU run on the same device
        c1.load_host(D)
        o1 = c1.order(); tr1 = c1.trace()
    same = bool((o == o1).all() and (tr == tr1).all() and (o2 == o1).all())
    ok &= same
    print(f"rank {rank}: {name} sharded == single: {same}", flush=True)
for n in (() if quick else (20000,)):
    c = fnn.Context(n, device=lr); c.connect_torch(); c.synth(1, 0.05)
    dist.barrier(); torch.cuda.synchronize(); t = time.perf_counter()
    o = c.order(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"rank {rank}: n={n} world={world} wall={dt:.3f}s", flush=True)
    c.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
