#!/bin/bash
# round-2 dev run 1: parity of the restructured tail + A/B timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_order.py tests/test_gpu_seqsum.py -m gpu -x -q > gpurun_out/d1_pytest_order.log 2>&1; echo "pytest order rc=$?"; tail -5 gpurun_out/d1_pytest_order.log
timeout 900 python -m pytest tests/test_gpu_modes.py -m gpu -x -q > gpurun_out/d1_pytest_modes.log 2>&1; echo "pytest modes rc=$?"; tail -5 gpurun_out/d1_pytest_modes.log
(
timeout 300 python tools/time_order.py --reps 2 20000
timeout 300 python tools/time_order.py --opt no_overlap=1 20000
timeout 300 python tools/time_order.py --opt force_exact_pick=1 20000
timeout 300 python tools/time_order.py --opt no_overlap=1 --opt force_exact_pick=1 20000
timeout 300 python tools/time_order.py 5000 10000
) > gpurun_out/d1_timing.log 2>&1
cat gpurun_out/d1_timing.log
