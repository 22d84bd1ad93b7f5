"""tools/time_csw.py [--variant V] n [n ...] : split-weight solve (seam B2) timing per n: CG iterations, us per iteration and the
roofline of one CG iteration (112 * npairs algorithmic bytes per iteration)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fastneighbornet_b200 as f
from fastneighbornet_b200 import synth
args = sys.argv[1:]
variant = "default"
if args and args[0] == "--variant":
    variant = args[1]; args = args[2:]
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
for n in [int(a) for a in args] or [200, 800]:
    D = synth.additive_noise_matrix(n, 1, 0.05)
    o = f.order(D)
    du = synth.upper_triangle(D)
    f.split_weights(o[:61] if False else o, du, constrained=False)   # warm-up: context, allocator
    t = time.time(); x, st = f.split_weights(o, du, variant=variant); dt = time.time() - t
    npairs = n * (n - 1) // 2
    us = 1e6 * dt / max(1, st["cg_iters"])
    gbs = 112.0 * npairs / (us * 1e-6) / 1e9
    print(f"n={n} variant={variant} wall={dt:.2f}s cg_iters={st['cg_iters']} cg_calls={st['cg_calls']} outer={st['outer']} inner={st['inner']} "
          f"launches={st['kernel_launches']} us_per_iter={us:.1f} alg={gbs:.0f} GB/s frac={gbs / peak:.3f} kept={(x > 1e-6).sum()}", flush=True)
