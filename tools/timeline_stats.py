"""tools/timeline_stats.py file.csv : per-kernel durations and gaps (microseconds) from an FNN_TIMELINE dump, bucketed by iteration."""
import sys
import numpy as np
a = np.genfromtxt(sys.argv[1], delimiter=",", names=True)
it = a["iter"]
def col(n): 
    v = a[n].astype(float); v[v < 0] = np.nan; return v / 1e3
c = {n: col(n) for n in a.dtype.names if n != "iter"}
nxt_scan0 = np.roll(c["scan0"], -1); nxt_scan0[-1] = np.nan
prev_scat1 = np.roll(c["scat1"], 1); prev_scat1[0] = np.nan
rows = {
    "scan": c["scan1"] - c["scan0"],
    "scan1->sel0 (boundary)": c["sel0"] - c["scan1"],
    "select decode (in k_rx_stage)": c["rx0"] - c["sel0"],
    "rx": c["rx1"] - c["rx0"],
    "rx1->pick0": c["pick0"] - c["rx1"],
    "pick": c["pick1"] - c["pick0"],
    "pick1->rows0": c["rows0"] - c["pick1"],
    "rows(blk0)": c["rows1"] - c["rows0"],
    "rows1->scat0": c["scat0"] - c["rows1"],
    "scatter": c["scat1"] - c["scat0"],
    "scat1->next scan0": nxt_scan0 - c["scat1"],
    "tail: scan1->next scan0": nxt_scan0 - c["scan1"],
    "iteration: scan0->next scan0": nxt_scan0 - c["scan0"],
    "chain0 - prev scat1": c["chain0"] - prev_scat1,
    "chain": c["chain1"] - c["chain0"],
    "patch": c["patch1"] - c["chain1"],
    "patch1 - scan1 (>0: join waits for the branch)": c["patch1"] - c["scan1"],
}
edges = [0, 500, 2000, 5000, 10000, 14000, 17000, 19000, 20000]
print("%-46s" % "us (median)" + "".join("%12s" % f"{lo}-{hi}" for lo, hi in zip(edges[:-1], edges[1:])))
for name, v in rows.items():
    out = []
    for lo, hi in zip(edges[:-1], edges[1:]):
        sel = (it >= lo) & (it < hi) & ~np.isnan(v)
        out.append(np.median(v[sel]) if sel.any() else np.nan)
    print("%-46s" % name + "".join("%12.1f" % x for x in out))
