#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_csw.py -m gpu -x -q -s > gpurun_out/d5_pytest_csw.log 2>&1; echo "pytest csw rc=$?"; grep -E "kept|passed|failed|Error" gpurun_out/d5_pytest_csw.log | tail -12
(
timeout 300 python tools/time_csw.py 200 400 800
FNN_CSW_GRID=148 timeout 300 python tools/time_csw.py 800
FNN_CSW_GRID=40 timeout 300 python tools/time_csw.py 800
timeout 600 python tools/time_csw.py 1500
) > gpurun_out/d5_csw_timing.log 2>&1
cat gpurun_out/d5_csw_timing.log
