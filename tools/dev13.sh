#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_csw.py tests/test_gpu_seqsum.py tests/test_gpu_cli.py -m gpu -x -q > gpurun_out/d13_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/d13_pytest.log
FNN_LIB=$PWD/fastneighbornet_b200/libfastnn_xsumtiming.so timeout 120 python tools/xsum_timing.py > gpurun_out/d13_xsum.log 2>&1; grep -E "chain 0|exact" gpurun_out/d13_xsum.log
(
timeout 300 python tools/time_csw.py 400 800
timeout 300 python tools/time_order.py --mode relaxed 20000
FNN_TIMELINE=0,20000,gpurun_out/d13_tl.csv timeout 300 python tools/time_order.py 20000
python tools/timeline_stats.py gpurun_out/d13_tl.csv | grep -E "median|^chain |^patch "
rm -f gpurun_out/d13_tl.csv
) > gpurun_out/d13.log 2>&1
cat gpurun_out/d13.log
