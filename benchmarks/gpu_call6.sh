#!/bin/bash
# round 1, session 4, third call: masked-epilogue prefetch, SpMM auto tuning + long-row threshold, other workloads
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_tc_gpu.py tests/test_gcond_gpu.py -q ) > gpurun_out/pytest_kernels.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_kernels.log; tail -4 gpurun_out/pytest_kernels.log
for thr in 128 64; do
  ( time timeout 300 python benchmarks/spmm_sweep.py --quick --variants --long-row $thr --out gpurun_out/spmm_var_$thr.json ) > gpurun_out/spmm_var_$thr.log 2>&1
  tail -3 gpurun_out/spmm_var_$thr.log | cut -c1-200
done
( time timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_arxiv.log 2>&1
tail -4 gpurun_out/bench_arxiv.log | cut -c1-300
for w in cora flickr reddit; do
  ( time timeout 500 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_$w.log 2>&1
  tail -4 gpurun_out/bench_$w.log | cut -c1-300
done
