"""tools/big_run.py <n> - Large-n run (BASELINE configs[4]: canonical, 100k taxa): torchrun or single process.  Prints time, algorithmic GB/s,
permutation invariants and a hash of the ordering (all ranks must agree)."""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastneighbornet_b200 as fnn
n = int(sys.argv[1])
rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
if world > 1:
    import torch.distributed as dist
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
c = fnn.Context(n, device=lr)
if world > 1:
    c.connect_torch()
t = time.perf_counter(); c.synth(5, 0.05); torch.cuda.synchronize(); tg = time.perf_counter() - t
if world > 1:
    dist.barrier()
t = time.perf_counter(); o = c.order(); dt = time.perf_counter() - t
st = c.stats()
h = hashlib.sha256(o.tobytes()).hexdigest()[:16]
ok = bool(o[0] == 0 and o[1] == 1 and (np.sort(o[1:]) == np.arange(1, n + 1)).all())
print(f"rank {rank}/{world}: n={n} synth {tg:.1f}s order {dt:.2f}s iters={st['iterations']} alg {st['scan_alg_bytes']/1e12:.1f} TB -> "
      f"{st['scan_alg_bytes']/dt/1e9:.0f} GB/s; permutation_ok={ok} sha={h}", flush=True)
if world > 1:
    hs = [None] * world
    dist.all_gather_object(hs, h)
    if rank == 0:
        print("all ranks agree:", len(set(hs)) == 1, flush=True)
    dist.barrier(); dist.destroy_process_group()
c.close()
