#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_order.py -m gpu -x -q -k "certified or production" > gpurun_out/d2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/d2_pytest.log
FNN_TIMELINE=0,20000,gpurun_out/d2_timeline.csv timeout 300 python tools/time_order.py 20000 > gpurun_out/d2_tl.log 2>&1; cat gpurun_out/d2_tl.log
python tools/timeline_stats.py gpurun_out/d2_timeline.csv > gpurun_out/d2_timeline_stats.txt 2>&1; cat gpurun_out/d2_timeline_stats.txt
gzip -f gpurun_out/d2_timeline.csv
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/d2_bench.json 2> gpurun_out/d2_bench.err; echo "bench rc=$?"; cat gpurun_out/d2_bench.json; tail -5 gpurun_out/d2_bench.err
