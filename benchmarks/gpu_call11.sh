#!/bin/bash
# final single-GPU pass of the session: full GPU test suite, smoke, bench (both arms), full SpMM sweep, ncu evidence
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log; tail -6 gpurun_out/pytest_gpu_final.log | cut -c1-300
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log | cut -c1-300
( time timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_arxiv_final.log 2>&1
tail -4 gpurun_out/bench_arxiv_final.log | cut -c1-300
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref_final.log 2>&1
tail -3 gpurun_out/bench_ref_final.log | cut -c1-300
( time timeout 600 python benchmarks/spmm_sweep.py --out gpurun_out/spmm_sweep_v3.json ) > gpurun_out/spmm_sweep_v3.log 2>&1
tail -3 gpurun_out/spmm_sweep_v3.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 3 -o gpurun_out/prof_real_gemm_v2 -f python profiles/capture_real_gemm.py > gpurun_out/capture_real_gemm_v2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 3 -o gpurun_out/prof_pge_gemm_v7 -f python profiles/capture_pge_gemm.py > gpurun_out/capture_pge_gemm_v7.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/launches_arxiv_v4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -12
