#!/bin/bash
# ncu --set full captures (one GPU): tall-skinny real-side GEMMs, PGE GEMMs, SpMM
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gcond_gpu.py tests/test_kernels_gpu.py -q -k "graph or adam or spmm" ) > gpurun_out/pytest_graph.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_graph.log; tail -15 gpurun_out/pytest_graph.log | cut -c1-300
for w in ogbn-arxiv cora; do
  ( time timeout 400 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_graph_$w.log 2>&1
  tail -4 gpurun_out/bench_graph_$w.log | cut -c1-250
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 3 -o gpurun_out/prof_real_gemm -f python profiles/capture_real_gemm.py > gpurun_out/capture_real_gemm.log 2>&1
tail -3 gpurun_out/capture_real_gemm.log
timeout 60 python profiles/capture_real_gemm.py 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 3 -o gpurun_out/prof_pge_gemm -f python profiles/capture_pge_gemm.py > gpurun_out/capture_pge_gemm.log 2>&1
tail -2 gpurun_out/capture_pge_gemm.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:spmm_wide -o gpurun_out/prof_spmm_v3 -f python profiles/capture_spmm.py > gpurun_out/capture_spmm_v3.log 2>&1
grep "^n=" gpurun_out/capture_spmm_v3.log
ls -la gpurun_out/*.ncu-rep
# sampler kernels at the Reddit shape (launch list restricted to the sampler's kernels)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ds::" -c 120 --csv --log-file gpurun_out/launches_reddit_sampler.csv python bench.py --workload reddit --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_reddit_under_ncu.log 2>&1
tail -2 gpurun_out/launches_reddit_sampler.csv | cut -c1-300
