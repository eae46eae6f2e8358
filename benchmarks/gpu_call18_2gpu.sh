#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/bench2_v8.log 2>&1
grep '^{"metric' gpurun_out/bench2_v8.log | tail -1 | cut -c1-200
( time timeout 200 python -m pytest tests/test_pge_sharded_gpu.py -q ) > gpurun_out/pytest_pge_sharded_2gpu_v2.log 2>&1; tail -2 gpurun_out/pytest_pge_sharded_2gpu_v2.log
