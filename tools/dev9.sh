#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_csw.py -m gpu -x -q > gpurun_out/d9_pytest_csw.log 2>&1; echo "pytest csw rc=$?"; tail -3 gpurun_out/d9_pytest_csw.log
(
FNN_CSW_PROF=1 timeout 300 python tools/time_csw.py 200 800
timeout 300 python tools/time_csw.py 400
FNN_CSW_GRID=74 timeout 300 python tools/time_csw.py 800
FNN_TIMELINE=0,20000,gpurun_out/d9_tl_relaxed.csv timeout 300 python tools/time_order.py --mode relaxed 20000
) > gpurun_out/d9.log 2>&1
cat gpurun_out/d9.log
python tools/timeline_stats.py gpurun_out/d9_tl_relaxed.csv > gpurun_out/d9_timeline_relaxed.txt 2>&1; cat gpurun_out/d9_timeline_relaxed.txt
rm -f gpurun_out/d9_tl_relaxed.csv
