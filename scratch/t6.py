import sys, time
sys.path.insert(0,'/root/repo')
import fastneighbornet_b200 as f
for n, mode, extra in [(20000,'relaxed',{}), (20000,'random_logn',{}), (5000,'random_n',{}), (5000,'random_nlogn',{})]:
    with f.Context(n, mode=mode, **extra) as c:
        c.synth(1, 0.05)
        t=time.time(); o=c.order(); dt=time.time()-t
        s=c.stats()
        print(f"n={n} mode={mode} wall={dt:.2f}s iters={s['iterations']} launches={s['kernel_launches']}", flush=True)
