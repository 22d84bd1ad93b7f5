#!/bin/bash
# what the driver runs at round end, in one call: GPU tests, smoke, the default bench line, a short reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; cat gpurun_out/final_bench.json | cut -c1-6000; tail -3 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --fit-n 2000,5000 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/final_bench_ref.json | cut -c1-3000
