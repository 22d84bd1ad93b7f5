#!/bin/bash
# N-GPU run: tools/dev12.sh N
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 tests/mg_check.py --quick > gpurun_out/d12_mgcheck_$N.log 2>&1; echo "mg_check rc=$?"; grep -c "sharded == single: True" gpurun_out/d12_mgcheck_$N.log; grep -E "False|Error|error" gpurun_out/d12_mgcheck_$N.log | head -5
timeout 900 $TR --master-port 29535 bench.py --gpus $N --steps 2 --warmup 2 > gpurun_out/d12_bench_$N.json 2> gpurun_out/d12_bench_$N.err; echo "bench rc=$?"; cut -c1-2200 gpurun_out/d12_bench_$N.json; tail -3 gpurun_out/d12_bench_$N.err
if [ "$N" = "8" ]; then
  timeout 600 $TR --master-port 29536 tools/big_run.py 100000 > gpurun_out/d12_big100k_$N.log 2>&1; echo "big rc=$?"; grep -E "rank 0|agree" gpurun_out/d12_big100k_$N.log
fi
