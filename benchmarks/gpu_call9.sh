#!/bin/bash
# epilogue rewrite check (LDS/STS staging, batched read-back), Reddit sampler kernel list
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gemm_tc_gpu.py tests/test_kernels_gpu.py -q -k "gemm or pge" ) > gpurun_out/pytest_gemm.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gemm.log; tail -5 gpurun_out/pytest_gemm.log | cut -c1-300
timeout 60 python profiles/capture_real_gemm.py 2>&1 | tail -1
( time timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_arxiv_epi.log 2>&1
tail -4 gpurun_out/bench_arxiv_epi.log | cut -c1-250
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sample_|mt_generate|pack_copy|pack_offsets" -c 60 --csv --log-file gpurun_out/launches_reddit_sampler.csv python bench.py --workload reddit --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_reddit_under_ncu.log 2>&1
tail -2 gpurun_out/launches_reddit_sampler.csv | cut -c1-300
