#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 5 --warmup 3 ) > gpurun_out/bench4_v8.log 2>&1
grep '^{"metric' gpurun_out/bench4_v8.log | tail -1 | cut -c1-200
