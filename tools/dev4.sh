#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_csw.py -m gpu -x -q -s > gpurun_out/d4_pytest_csw.log 2>&1; echo "pytest csw rc=$?"; grep -E "^n=|passed|failed|Error|error" gpurun_out/d4_pytest_csw.log | tail -20
(
timeout 300 python tools/time_csw.py 200 400 800
timeout 200 python tools/time_csw.py --variant graph 200
FNN_CSW_GRID=148 timeout 300 python tools/time_csw.py 800
FNN_CSW_GRID=32 timeout 300 python tools/time_csw.py 800
) > gpurun_out/d4_csw_timing.log 2>&1
cat gpurun_out/d4_csw_timing.log
