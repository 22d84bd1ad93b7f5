#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/time_csw.py 400"
export FNN_CSW_ITERS_PER_LAUNCH=3000
$CMD > gpurun_out/d8_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_cg_persistent -s 20 -c 1 -o gpurun_out/prof_cg $CMD > gpurun_out/d8_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/d8_plain.log; tail -5 gpurun_out/d8_ncu.log
unset FNN_CSW_ITERS_PER_LAUNCH
timeout 600 python -m pytest tests/test_gpu_order.py -m gpu -x -q > gpurun_out/d8_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/d8_pytest.log
FNN_TIMELINE=0,20000,gpurun_out/d8_timeline.csv timeout 300 python tools/time_order.py 20000 > gpurun_out/d8_tl.log 2>&1; cat gpurun_out/d8_tl.log
python tools/timeline_stats.py gpurun_out/d8_timeline.csv > gpurun_out/d8_timeline_stats.txt 2>&1; cat gpurun_out/d8_timeline_stats.txt
rm -f gpurun_out/d8_timeline.csv
timeout 300 python tools/time_order.py --reps 2 20000
