"""tools/config1_run.py [n] : BASELINE configs[1] - canonical full network (ordering + CircularSplitWeights) on n taxa (default 5000), one B200.
Prints progress on stderr (FNN_CSW_PROGRESS) and one JSON summary; size-independent checks: the ordering is a permutation, the
weights are non-negative, and the fitted distances A x reproduce the input up to the residual the NNLS leaves."""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("FNN_CSW_PROGRESS", "60")
import numpy as np
import fastneighbornet_b200 as f
from fastneighbornet_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
out = sys.argv[2] if len(sys.argv) > 2 else None
D = synth.additive_noise_matrix(n, 2, 0.05)
f.order(synth.additive_noise_matrix(64, 1, 0.05))          # context / allocator warm-up
t = time.time(); o = f.order(D); t_order = time.time() - t
f.release_cache()
du = synth.upper_triangle(D)
t = time.time(); x, st = f.split_weights(o, du); t_sw = time.time() - t
npairs = n * (n - 1) // 2
kept = int((x > 1e-6).sum())
taxa = np.concatenate([[o[n]], o[1:n]]) - 1
d_pos = D[np.ix_(taxa, taxa)][np.triu_indices(n, 1)]
ax = f.csw_matvec("ab", x, n)
res = {"config": f"BASELINE configs[1]: Canonical full network (order + CircularSplitWeights), {n} taxa, additive tree + 5% noise, 1 B200",
       "n_taxa": n, "ordering_s": t_order, "ordering_sha256": hashlib.sha256(o.tobytes()).hexdigest()[:16],
       "permutation_ok": bool(o[0] == 0 and o[1] == 1 and (np.sort(o[1:]) == np.arange(1, n + 1)).all()),
       "split_weights_s": t_sw, "cg_iterations": st["cg_iters"], "cg_solves": st["cg_calls"], "outer": st["outer"], "inner": st["inner"],
       "kernel_launches": st["kernel_launches"], "us_per_cg_iteration": 1e6 * t_sw / max(1, st["cg_iters"]),
       "alg_gbs": 112.0 * npairs * st["cg_iters"] / t_sw / 1e9, "kept_splits": kept, "min_weight": float(x.min()),
       "fit_rel_residual": float(np.linalg.norm(ax - d_pos) / np.linalg.norm(d_pos)), "total_s": t_order + t_sw}
line = json.dumps(res)
print(line, flush=True)
if out:
    open(out, "w").write(line + "\n")
