#!/bin/bash
# two GPUs: row-sharded PGE over NCCL (unit test + class-sharded bench with it on and off)
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_pge_sharded_gpu.py tests/test_row_partition_gpu.py -q ) > gpurun_out/pytest_pge_sharded_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_pge_sharded_2gpu.log; tail -12 gpurun_out/pytest_pge_sharded_2gpu.log | cut -c1-300
( time timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "sharded" ) > gpurun_out/pytest_pge_pieces.log 2>&1
tail -5 gpurun_out/pytest_pge_pieces.log | cut -c1-300
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/bench2_pge_sharded.log 2>&1
tail -5 gpurun_out/bench2_pge_sharded.log | cut -c1-500
