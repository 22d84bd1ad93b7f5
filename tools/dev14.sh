#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_order.py tests/test_gpu_modes.py -m gpu -x -q > gpurun_out/d14_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/d14_pytest.log
(
FNN_TIMELINE=0,20000,gpurun_out/d14_tl.csv timeout 300 python tools/time_order.py 20000
python tools/timeline_stats.py gpurun_out/d14_tl.csv
rm -f gpurun_out/d14_tl.csv
timeout 300 python tools/time_order.py --reps 2 20000
timeout 300 python tools/time_order.py --mode relaxed 20000
FNN_CSW_ABORT_AFTER=30000 timeout 600 python tools/time_csw.py 5000
) > gpurun_out/d14.log 2>&1
cat gpurun_out/d14.log | tail -40
