#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for cfg in "1 2" "1 1" "0 4"; do
  set -- $cfg
  GS_TC_SHRINK_BN=$1 GS_TC_MIN_KB=$2 timeout 300 python benchmarks/gemm_shapes.py --out gpurun_out/gemm_shapes_s$1_k$2.json > gpurun_out/gemm_shapes_s$1_k$2.log 2>&1
  echo "shrink=$1 min_kb=$2: $(tail -1 gpurun_out/gemm_shapes_s$1_k$2.log)"
done
for cfg in "8 4" "4 4" "2 4" "4 8" "2 8"; do
  set -- $cfg
  GS_SPMM_WPB=$1 GS_SPMM_UNR=$2 timeout 300 python benchmarks/spmm_sweep.py --quick --out gpurun_out/spmm_quick_w$1_u$2.json > gpurun_out/spmm_quick_w$1_u$2.log 2>&1
  echo "wpb=$1 unr=$2: $(grep -o '"ms": [0-9.]*' gpurun_out/spmm_quick_w$1_u$2.log | tr '\n' ' ')"
done
( time timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_arxiv_v3.log 2>&1
tail -4 gpurun_out/bench_arxiv_v3.log | cut -c1-400
