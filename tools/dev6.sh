#!/bin/bash
# 2-GPU validation: sharded scan + mailbox exchange merged in k_rx_stage, replicated tail, forked chain/patch
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29533 tests/mg_check.py > gpurun_out/d6_mgcheck.log 2>&1; echo "mg_check rc=$?"; grep -E "sharded|wall" gpurun_out/d6_mgcheck.log
FNN_TIMELINE=0,20000,gpurun_out/d6_tl.csv timeout 300 $TR --master-port 29534 tools/big_run.py 20000 > gpurun_out/d6_tl.log 2>&1; grep -E "rank|agree" gpurun_out/d6_tl.log
python tools/timeline_stats.py gpurun_out/d6_tl.csv.r0 > gpurun_out/d6_timeline_stats_rank0.txt 2>&1; cat gpurun_out/d6_timeline_stats_rank0.txt
rm -f gpurun_out/d6_tl.csv.r0 gpurun_out/d6_tl.csv.r1
timeout 900 $TR --master-port 29535 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/d6_bench2.json 2> gpurun_out/d6_bench2.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/d6_bench2.json; tail -3 gpurun_out/d6_bench2.err
