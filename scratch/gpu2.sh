python scratch/t2.py 20000 2>&1 | grep -E "tma|identical"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_scan_tma -s 20 -c 2 -o gpurun_out/prof_scan_tma $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/plain.log | cut -c1-600
