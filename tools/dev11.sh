#!/bin/bash
mkdir -p gpurun_out
FNN_LIB=$PWD/fastneighbornet_b200/libfastnn_xsumtiming.so timeout 120 python tools/xsum_timing.py > gpurun_out/d11_xsum.log 2>&1; cat gpurun_out/d11_xsum.log | head -30
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/d11_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/d11_pytest_all.log
(
timeout 300 python tools/time_order.py --mode relaxed --reps 2 20000
timeout 300 python tools/time_order.py --mode random_nlogn 10000
) > gpurun_out/d11.log 2>&1
cat gpurun_out/d11.log
