#!/bin/bash
# launch list of the bench command (B200_PROFILING.md): every launch with its device time, cold-cache and serialised
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-configs"
$CMD > gpurun_out/ncu_ll_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 7000 -c 700 --csv --log-file gpurun_out/r2_launches_bench_n20000.csv $CMD > gpurun_out/ncu_ll.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_ll_plain.log | cut -c1-300; tail -3 gpurun_out/ncu_ll.log
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r2_launches_bench_n20000.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].split("::")[-1]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:24s} launches {c:5d}  avg {t / c / 1e3:8.2f} us  share {100 * t / tot:5.1f} %")
PY
