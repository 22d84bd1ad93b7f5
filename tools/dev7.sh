#!/bin/bash
mkdir -p gpurun_out
(
FNN_CSW_PROF=1 timeout 300 python tools/time_csw.py 200 800
FNN_CSW_PROF=1 FNN_CSW_GRID=148 timeout 300 python tools/time_csw.py 800
timeout 300 python tools/time_order.py --mode relaxed 20000
timeout 300 python tools/time_order.py --mode relaxed --opt additive=1 --eps 0 20000
timeout 300 python tools/time_order.py --mode random_logn 20000
) > gpurun_out/d7.log 2>&1
cat gpurun_out/d7.log
