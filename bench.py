#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the Neighbor-Net hot path.

A "step" is one full canonical Neighbor-Net ordering (NetMakerOriginal.runNeighborNet,
/root/reference NetMakerOriginal.java:129) of a synthetic n-taxon matrix (additive tree + 5 %
noise).  metric/unit: algorithmic HBM GB/s of the job = the selection scan's
algorithmic bytes (8 B per cross-cluster matrix entry per iteration; summed exactly
on the device) divided by the wall time of the whole ordering, so it is comparable across n and
between the GPU and the CPU arms; `ms_per_step` is the ordering's wall time.

  value     inputs resident in HBM (device matrix restored from a pristine device copy each step)
  e2e       the same through the reference-facing one-shot C-ABI call fnn_order() with a pinned
            HOST matrix: cudaMalloc + H2D of n*n*8 bytes + ordering + D2H of the ordering, all timed
  roofline  the selection kernel k_scan: algorithmic bytes per launch / CUDA-event launch time,
            sampled every 16th iteration inside a dedicated profiled run of the same workload
  cpu_baseline / --impl reference: the CPU restatement of the reference (oracle/, "port": the JAR
            cannot run here - no JVM) on a bounded sample (smaller n), all host threads.

N>1: one process per GPU, ONE job: the selection scan (the Theta(n^3) part) is sharded across the
ranks, the per-rank (Q,i,j) partial min-locs are exchanged through peer-mapped mailboxes over NVLink
inside the kernels, the Theta(n^2) update is replicated -> scaling is "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "neighbornet_canonical_order_algorithmic_hbm_throughput"
UNIT = "GB/s"


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_reference_run(n, seed, threads):
    """One canonical ordering on the CPU restatement; returns (seconds, algorithmic bytes)."""
    import oracle
    from fastneighbornet_b200 import synth
    D = synth.additive_noise_matrix(n, seed, 0.05)
    t0 = time.perf_counter()
    _, _, info = oracle.order(D, mode="canonical", want_trace=False, threads=threads)
    dt = time.perf_counter() - t0
    return dt, info["alg_bytes"]


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (restatement; the JAR needs a JVM that this
    image does not have) on a bounded sample of the workload, all host threads, rank 0 only."""
    if rank != 0:
        return
    threads = host_threads()
    n = args.ref_n
    for i in range(args.warmup):
        cpu_reference_run(n, 100 + i, threads)
    tot_t, tot_b = 0.0, 0.0
    for i in range(args.steps):
        dt, b = cpu_reference_run(n, 1 + i, threads)
        tot_t += dt
        tot_b += b
    val = tot_b / tot_t / 1e9
    sample = f"canonical ordering, n={n} (same generator, eps=0.05), {threads} threads, NeighborNetCanonical thread partition"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        "workload": f"canonical Neighbor-Net ordering (-mode Canonical -order), n={args.n} taxa, synthetic additive tree + 5% noise "
                    f"(BASELINE metric 'n=20k'; fits one GPU: {args.n * args.n * 8 / 1e9:.1f} GB fp64 matrix)",
        "n_taxa": args.n, "mode": "canonical", "eps": 0.05,
        "parallelism": "single GPU" if world == 1 else f"{world} GPUs: scan sharded by tile, P2P min-loc mailbox exchange, replicated update",
        "l2": f"input matrix {args.n * args.n * 8 / 1e6:.0f} MB >> 126 MB L2; no explicit flush",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--ref-n", type=int, default=2000, help="taxa of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import fastneighbornet_b200 as fnn

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.n
    seed = 1   # every rank holds the same matrix (one job)
    ctx = fnn.Context(n, device=local_rank, use_graph=1)
    if world > 1:
        ctx.connect_torch()
    # pristine device copy of the input (torch owns it: plumbing, not the product)
    ctx.synth(seed, 0.05)
    dptr, ld = ctx.matrix_ptr()
    pristine = torch.empty((n, ld), dtype=torch.float64, device=f"cuda:{local_rank}")
    # view of the library's device matrix through __cuda_array_interface__ (zero copy)

    class _View:
        def __init__(self, ptr, shape):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (ptr, False), "version": 3}

    view = torch.as_tensor(_View(dptr, (n, ld)), device=f"cuda:{local_rank}")
    pristine.copy_(view)
    torch.cuda.synchronize()

    def step():
        ctx.load_device(pristine.data_ptr(), ld)
        return ctx.order()

    ordering0 = None
    for _ in range(args.warmup):
        ordering0 = step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t0 = time.perf_counter()
    dev_ms, launches, alg_bytes = 0.0, 0, 0.0
    for _ in range(args.steps):
        o = step()
        st = ctx.stats()
        dev_ms += st["order_ms"]
        launches += st["kernel_launches"]
        alg_bytes += st["scan_alg_bytes"]
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    assert (o == ordering0).all(), "ordering changed between identical steps"
    # max over ranks of the elapsed time; bytes summed over ranks
    if world > 1:
        t = torch.tensor([elapsed, dev_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, dev_ms = float(t[0]), float(t[1])
        b = torch.tensor([float(launches)], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        launches = int(b[0])   # alg_bytes is the ONE job's (identical on every rank)
    value = alg_bytes / elapsed / 1e9

    # ---- e2e: one-shot C-ABI call with a pinned HOST matrix
    e2e = None
    if not args.no_e2e:
        host = torch.empty((n, n), dtype=torch.float64).pin_memory()
        host.copy_(pristine[:, :n])  # the ctx matrix itself was consumed by the runs above
        torch.cuda.synchronize()
        Dh = host.numpy()

        def e2e_step():
            if world == 1:
                return fnn.order(Dh, device=local_rank)   # the one-shot reference-facing seam (fnn_order)
            ctx.load_host(Dh)                             # N>1: same public API on the wired context
            return ctx.order()

        if world == 1:
            ctx.close()
        e2e_step()  # warm-up (allocator, graph instantiation)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            oe = e2e_step()
        barrier()
        e2e_elapsed = time.perf_counter() - t0
        assert (oe == ordering0).all(), "e2e ordering differs from the device-resident run"
        if world > 1:
            t = torch.tensor([e2e_elapsed], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_elapsed = float(t[0])
        e2e = {"value": alg_bytes / e2e_elapsed / 1e9, "unit": UNIT, "h2d_bytes_per_step": n * n * 8 * world,
               "d2h_bytes_per_step": (n + 1) * 4 * world, "ms_per_step": 1e3 * e2e_elapsed / args.steps}
        if world == 1:
            ctx = fnn.Context(n, device=local_rank, use_graph=1)

    # ---- roofline of the dominant kernel (k_scan), rank 0 only
    roofline, cpu_baseline = None, None
    if rank == 0:
        pctx = fnn.Context(n, device=local_rank, profile_every=16)
        pctx.load_device(pristine.data_ptr(), ld)
        pctx.order()
        ps = pctx.stats()
        pctx.close()
        peak, peak_src = measured_peak()
        achieved = ps["prof_scan_bytes"] / (ps["prof_scan_ms"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_scan_tma", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "frac_of_nominal_8000": achieved / 8000.0,
                    "traffic": None, "peak_source": peak_src, "samples": ps["prof_scan_samples"],
                    "avg_launch_ms": ps["prof_scan_ms"] / max(1, ps["prof_scan_samples"]),
                    "alg_bytes_per_launch": ps["prof_scan_bytes"] / max(1, ps["prof_scan_samples"]),
                    "scan_share_of_step": ((ps["prof_scan_ms"] * 16) / (1e3 * elapsed / args.steps)) if world == 1 else None}
        if world > 1:
            roofline["note"] = "kernel profiled on rank 0 as a single-GPU run of the same workload (each rank launches it on 1/world of the tiles)"
        tfile = os.path.join(ROOT, "profiles", "scan_traffic.json")
        if os.path.exists(tfile):
            try:
                with open(tfile) as f:
                    roofline["traffic"] = json.load(f).get("traffic_bytes_per_launch")
            except Exception:
                pass
        if not args.no_cpu_baseline and world == 1:   # reported at N=1 only
            threads = host_threads()
            dt, b = cpu_reference_run(args.ref_n, 1, threads)
            cpu_baseline = {"value": b / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"one canonical ordering at n={args.ref_n} (same generator), {dt:.1f} s on {threads} threads"}
            n1 = min(args.ref_n, 1500)   # the reference's -threads 1 figure on a smaller sample
            dt1, b1 = cpu_reference_run(n1, 1, 1)
            cpu_baseline["value_1thread"] = b1 / dt1 / 1e9
            cpu_baseline["sample_1thread"] = f"n={n1}, {dt1:.1f} s on 1 thread"
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "device_ms_per_step": dev_ms / args.steps, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
