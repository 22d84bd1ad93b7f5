#!/bin/bash
# tools/sanitize.sh <memcheck|racecheck|synccheck> [small|tiny] : ONE compute-sanitizer tool per gpurun call
TOOL=$1; SIZE=${2:-small}
mkdir -p gpurun_out
python tools/sanitize_case.py $SIZE > gpurun_out/san_plain_$TOOL.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/san_plain_$TOOL.log; exit 1; }
tail -1 gpurun_out/san_plain_$TOOL.log
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_case.py $SIZE > gpurun_out/san_$TOOL.log 2>&1
echo "sanitizer rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ALL OK|FAILED|MISMATCH" gpurun_out/san_$TOOL.log | tail -8
grep -E "=========" gpurun_out/san_$TOOL.log | head -40
