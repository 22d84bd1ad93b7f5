"""tools/time_order.py [--mode M] n [n ...] : wall time of one ordering per n (device-generated input)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastneighbornet_b200 as f
args = sys.argv[1:]
mode = "canonical"
if args and args[0] == "--mode":
    mode = args[1]; args = args[2:]
for n in [int(a) for a in args] or [5000, 20000]:
    with f.Context(n, mode=mode) as c:
        c.synth(1, 0.05)
        t = time.time(); c.order(); dt = time.time() - t
        s = c.stats()
        print(f"n={n} mode={mode} wall={dt:.3f}s iterations={s['iterations']} launches={s['kernel_launches']} "
              f"alg={s['scan_alg_bytes']/dt/1e9:.0f} GB/s", flush=True)
