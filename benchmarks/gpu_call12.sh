#!/bin/bash
# evaluator GPU parity + refreshed bench lines of the three other shapes (Cora with its CPU baseline)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_evaluator.py tests/test_gcond_gpu.py -q -m gpu ) > gpurun_out/pytest_eval_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_eval_gpu.log; tail -8 gpurun_out/pytest_eval_gpu.log | cut -c1-300
( time timeout 500 python bench.py --workload cora --steps 5 --warmup 3 ) > gpurun_out/bench_cora_v5.log 2>&1
tail -3 gpurun_out/bench_cora_v5.log | cut -c1-250
for w in flickr reddit; do
  ( time timeout 500 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_${w}_v5.log 2>&1
  tail -3 gpurun_out/bench_${w}_v5.log | cut -c1-250
done
