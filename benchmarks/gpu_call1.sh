#!/bin/bash
# tests + bench + launch list + ncu full captures (one GPU)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_arxiv.log 2>&1
( time timeout 300 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_arxiv.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_wide_kernel -o gpurun_out/prof_spmm -f python profiles/capture_spmm.py > gpurun_out/capture_spmm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 3 -o gpurun_out/prof_pge_gemm -f python profiles/capture_pge_gemm.py > gpurun_out/capture_gemm.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_arxiv.log
