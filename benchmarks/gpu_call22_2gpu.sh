#!/bin/bash
mkdir -p gpurun_out
( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/bench2_final.log 2>&1
grep '^{"metric' gpurun_out/bench2_final.log | tail -1 | cut -c1-200; tail -3 gpurun_out/bench2_final.log | cut -c1-200
