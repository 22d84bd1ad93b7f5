set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches_b.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_scan_tma -s 20 -c 2 -o gpurun_out/prof_scan_tma_b $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
