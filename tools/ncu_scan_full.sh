#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-configs"
$CMD > gpurun_out/ncu_sf_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_scan_tma -s 20 -c 2 -o gpurun_out/prof_scan_tma_r2 $CMD > gpurun_out/ncu_sf.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_sf.log
ncu -i gpurun_out/prof_scan_tma_r2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_k_scan_tma_raw.csv 2>/dev/null; wc -c gpurun_out/r2_ncu_full_k_scan_tma_raw.csv
