#!/bin/bash
# round 1, session 4, second call: GEMM regression fix check, SpMM v1 tuning sweep, shrink-BN A/B
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_tc_gpu.py -q ) > gpurun_out/pytest_kernels.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_kernels.log; tail -4 gpurun_out/pytest_kernels.log
for thr in 512 128; do
  ( time timeout 300 python benchmarks/spmm_sweep.py --quick --variants --long-row $thr --out gpurun_out/spmm_var_$thr.json ) > gpurun_out/spmm_var_$thr.log 2>&1
  tail -3 gpurun_out/spmm_var_$thr.log | cut -c1-200
done
( time timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_arxiv_shrink1.log 2>&1
tail -4 gpurun_out/bench_arxiv_shrink1.log | cut -c1-300
( time GS_TC_SHRINK_BN=0 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_arxiv_shrink0.log 2>&1
tail -4 gpurun_out/bench_arxiv_shrink0.log | cut -c1-300
for s in 1 0; do
  GS_TC_SHRINK_BN=$s timeout 300 python benchmarks/gemm_shapes.py --out gpurun_out/gemm_shapes_s$s.json > gpurun_out/gemm_shapes_s$s.log 2>&1
  echo "shrink=$s: $(tail -1 gpurun_out/gemm_shapes_s$s.log)"
done
