"""tools/sanitize_case.py [small|tiny] : the smallest runs that touch every kernel family, for compute-sanitizer
(one tool per gpurun call, B200_PROFILING.md).  Results are checked against the oracle so a silent mis-execution under the
tool is noticed too."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fastneighbornet_b200 as f
import oracle
from fastneighbornet_b200 import synth
tiny = len(sys.argv) > 1 and sys.argv[1] == "tiny"
nc, nr, ns = (150, 90, 16) if tiny else (300, 200, 30)
ok = True
def check(name, cond):
    global ok
    ok &= bool(cond)
    print(f"{name}: {'ok' if cond else 'MISMATCH'}", flush=True)
D = synth.additive_noise_matrix(nc, 3, 0.05)
o_ref, tr_ref, _ = oracle.order(D)
for opts in ({"record_trace": 1}, {}, {"no_overlap": 1}, {"use_graph": 0}):
    with f.Context(nc, **opts) as c:
        c.load_host(D); o = c.order()
        check(f"canonical n={nc} {opts}", (o == o_ref).all() and (not opts.get("record_trace") or (c.trace() == tr_ref).all()))
D = synth.additive_noise_matrix(nr, 4, 0.05)
for mode, extra in (("relaxed", {}), ("relaxed", {"additive": 1}), ("random_n", {}), ("random_logn", {})):
    if mode == "relaxed" and extra and not tiny:
        Dm = synth.additive_noise_matrix(60, 4, 0.05)
    else:
        Dm = D if not extra else synth.additive_noise_matrix(40, 4, 0.05)
    o_ref, _, _ = oracle.order(Dm, mode=mode, seed=5, fallback=8, additive=bool(extra.get("additive")), want_trace=False)
    with f.Context(Dm.shape[0], mode=mode, seed=5, canonical_fallback=8, **extra) as c:
        c.load_host(Dm); o = c.order()
    check(f"{mode} {extra} n={Dm.shape[0]}", (o == o_ref).all())
D = synth.additive_noise_matrix(ns, 2, 0.05)
o = f.order(D); du = synth.upper_triangle(D)
d_pos = oracle.setup_d(o, du)
x1, s1 = oracle.l1_split_weights(ns, d_pos)
x0, s0 = oracle.split_weights(ns, d_pos)
for variant, xr, sr in (("default", x1, s1), ("graph", x1, s1), ("literal", x0, s0)):
    x, st = f.split_weights(o, du, variant=variant)
    check(f"split weights {variant} n={ns} ({st['cg_iters']} CG iterations)", (x == xr).all() and st["cg_iters"] == sr["cg_iters"])
f.release_cache()
print("ALL OK" if ok else "FAILED", flush=True)
sys.exit(0 if ok else 1)
