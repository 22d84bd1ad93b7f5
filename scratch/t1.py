import time, sys
sys.path.insert(0,'/root/repo')
import fastneighbornet_b200 as f
for n in [2000, 5000, 10000, 20000]:
    for prof in [0, 16]:
        with f.Context(n, profile_every=prof) as c:
            c.synth(1, 0.05)
            t=time.time(); o=c.order(); dt=time.time()-t
            s=c.stats()
            line = f"n={n} prof={prof} wall={dt:.3f}s order_ms={s['order_ms']:.1f} iters={s['iterations']} launches={s['kernel_launches']}"
            if prof: line += f" scan: {s['prof_scan_samples']} samples {s['prof_scan_ms']:.2f} ms, {s['prof_scan_bytes']/s['prof_scan_ms']/1e6:.1f} GB/s; est total scan share"
            print(line, flush=True)
