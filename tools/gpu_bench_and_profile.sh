#!/bin/bash
# smoke, default bench line, then the two ncu passes of B200_PROFILING.md (launch list + full capture of k_scan_tma)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"; cat gpurun_out/bench_d.json | cut -c1-2500
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches_d.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_scan_tma -s 20 -c 2 -o gpurun_out/prof_scan_tma_d $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
