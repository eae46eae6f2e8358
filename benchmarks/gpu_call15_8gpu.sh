#!/bin/bash
# one 8-GPU box: the driver's scaling run at the arxiv shape (N = 2, 4, 8; N = 1 for the same box as reference point)
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/smi8.txt
( time timeout 200 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/scale_n1.log 2>&1
tail -1 gpurun_out/scale_n1.log | cut -c1-160
for n in 2 4 8; do
  ( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 ) > gpurun_out/scale_n$n.log 2>&1
  grep '^{"metric' gpurun_out/scale_n$n.log | tail -1 | cut -c1-160
done
