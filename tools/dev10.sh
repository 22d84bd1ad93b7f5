#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py tests/test_gpu_order.py -m gpu -x -q > gpurun_out/d10_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/d10_pytest.log
(
FNN_TIMELINE=0,20000,gpurun_out/d10_tl.csv timeout 300 python tools/time_order.py 20000
python tools/timeline_stats.py gpurun_out/d10_tl.csv | grep -E "median|pick|tail|iteration"
FNN_TIMELINE=0,20000,gpurun_out/d10_tlr.csv timeout 300 python tools/time_order.py --mode relaxed 20000
python tools/timeline_stats.py gpurun_out/d10_tlr.csv | grep -E "median|^scan |tail|iteration"
rm -f gpurun_out/d10_tl.csv gpurun_out/d10_tlr.csv
timeout 300 python tools/time_order.py --reps 2 20000
timeout 300 python tools/time_order.py --mode relaxed --reps 2 20000
timeout 900 python tools/time_csw.py 1500 3000
) > gpurun_out/d10.log 2>&1
cat gpurun_out/d10.log
