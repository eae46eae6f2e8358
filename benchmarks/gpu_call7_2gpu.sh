#!/bin/bash
# two GPUs: NCCL row-partitioned SpMM test, class-sharded bench at the arxiv and Reddit shapes
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.txt
( time timeout 300 python -m pytest tests/test_row_partition_gpu.py -q ) > gpurun_out/pytest_rowpart_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_rowpart_2gpu.log; tail -4 gpurun_out/pytest_rowpart_2gpu.log
for w in ogbn-arxiv reddit; do
  ( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --workload $w ) > gpurun_out/bench2_$w.log 2>&1
  tail -5 gpurun_out/bench2_$w.log | cut -c1-400
done
