CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 56000 -c 560 --csv --log-file gpurun_out/launches_c.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu rc=$?"
