#!/bin/bash
# tools/scale_bench.sh N : bench.py on N GPUs of one box (what the driver's scaling step runs), output in gpurun_out/
N=$1
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
fi
echo "rc=$?"; tail -c 1500 gpurun_out/scale_$N.json; tail -3 gpurun_out/scale_$N.err | cut -c1-300
