#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_order.py tests/test_gpu_modes.py -m gpu -x -q > gpurun_out/d3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/d3_pytest.log
FNN_TIMELINE=0,20000,gpurun_out/d3_timeline.csv timeout 300 python tools/time_order.py 20000 > gpurun_out/d3_tl.log 2>&1; cat gpurun_out/d3_tl.log
python tools/timeline_stats.py gpurun_out/d3_timeline.csv > gpurun_out/d3_timeline_stats.txt 2>&1; cat gpurun_out/d3_timeline_stats.txt
rm -f gpurun_out/d3_timeline.csv
timeout 300 python tools/time_order.py --reps 2 20000 10000 40000
timeout 300 python tools/time_order.py --mode relaxed 20000
