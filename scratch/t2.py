import time, sys
sys.path.insert(0,'/root/repo')
import fastneighbornet_b200 as f
import numpy as np
ns = [int(x) for x in sys.argv[1:]] or [5000, 20000]
for n in ns:
    outs = []
    for impl in [1, 0]:
        for prof in [0, 16]:
            o_ = f.default_opts()
            c = f.Context(n, profile_every=prof)
            c.close()
            opts = dict(profile_every=prof)
            c = f.Context.__new__(f.Context)
            c.n = n; c.opts = f.default_opts(**opts); c.opts.reserved[0] = impl
            import ctypes
            c._h = ctypes.c_void_p()
            f.api._check(f.lib().fnn_ctx_create(ctypes.byref(c.opts), n, ctypes.byref(c._h)))
            c.synth(1, 0.05)
            t=time.time(); o=c.order(); dt=time.time()-t
            s=c.stats(); outs.append(o)
            line = f"n={n} impl={'tma' if impl==0 else 'reg'} prof={prof} wall={dt:.3f}s order_ms={s['order_ms']:.1f} iters={s['iterations']}"
            if prof: line += f" scan: {s['prof_scan_ms']*16/1e3:.2f}s est total, {s['prof_scan_bytes']/s['prof_scan_ms']/1e6:.1f} GB/s"
            print(line, flush=True)
            c.close()
    assert all((o == outs[0]).all() for o in outs), "orderings differ between scan implementations"
    print("orderings identical across impls")
