#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu_v3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v3.log; tail -5 gpurun_out/pytest_gpu_v3.log | cut -c1-300
for w in ogbn-arxiv reddit flickr cora; do
  ( time timeout 500 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_${w}_v8.log 2>&1
  grep '^{"metric' gpurun_out/bench_${w}_v8.log | tail -1 | cut -c1-160
done
