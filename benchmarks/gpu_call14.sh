#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_gcond_gpu.py tests/test_evaluator.py -q -m gpu ) > gpurun_out/pytest_splitk.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_splitk.log; tail -6 gpurun_out/pytest_splitk.log | cut -c1-300
for w in cora reddit; do
  ( time timeout 500 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_${w}_v6.log 2>&1
  tail -3 gpurun_out/bench_${w}_v6.log | cut -c1-200
done
timeout 300 python benchmarks/gemm_shapes.py --workload cora --out gpurun_out/gemm_shapes_cora_v2.json > gpurun_out/gemm_shapes_cora_v2.log 2>&1; tail -1 gpurun_out/gemm_shapes_cora_v2.log
