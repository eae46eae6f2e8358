#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "spmm" ) > gpurun_out/pytest_spmm.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_spmm.log
tail -5 gpurun_out/pytest_spmm.log
( time timeout 300 python benchmarks/spmm_sweep.py --quick --out gpurun_out/spmm_sweep_quick.json ) > gpurun_out/spmm_sweep_quick.log 2>&1
tail -6 gpurun_out/spmm_sweep_quick.log | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_wide_kernel -o gpurun_out/prof_spmm_v2 -f python profiles/capture_spmm.py > gpurun_out/capture_spmm_v2.log 2>&1
( time timeout 300 python benchmarks/gemm_shapes.py ) > gpurun_out/gemm_shapes.log 2>&1
tail -3 gpurun_out/gemm_shapes.log
( time timeout 500 python bench.py --workload reddit --steps 2 --warmup 2 --no-cpu-baseline ) > gpurun_out/bench_reddit.log 2>&1
tail -4 gpurun_out/bench_reddit.log | cut -c1-3000
