#!/bin/bash
# round 1, session 4: parity tests, SpMM tuning sweep, arxiv + Reddit bench, launch list, ncu full of the SpMM (one GPU)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
# the new SpMM generation first, in its own process: if it is broken the rest of the call runs on the v1 kernel
if timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "spmm" > gpurun_out/pytest_spmm.log 2>&1; then
  echo "spmm tests green (v2)"
else
  echo "spmm tests FAILED on v2 -> GS_SPMM_IMPL=1 for the rest"; tail -30 gpurun_out/pytest_spmm.log; export GS_SPMM_IMPL=1
fi
( time timeout 1000 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
( time timeout 400 python benchmarks/spmm_sweep.py --quick --variants --out gpurun_out/spmm_sweep_variants.json ) > gpurun_out/spmm_sweep_variants.log 2>&1
grep -o '"n": [0-9]*\|"F": [0-9]*\|"variants_ms": {[^}]*}' gpurun_out/spmm_sweep_variants.log | paste - - - | cut -c1-420
( time timeout 600 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_arxiv.log 2>&1
tail -4 gpurun_out/bench_arxiv.log | cut -c1-600
( time timeout 700 python bench.py --workload reddit --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_reddit.log 2>&1
tail -4 gpurun_out/bench_reddit.log | cut -c1-2500
timeout 500 ncu --set full --clock-control none --import-source on -k regex:spmm_wide -o gpurun_out/prof_spmm_v2 -f python profiles/capture_spmm.py > gpurun_out/capture_spmm_v2.log 2>&1
tail -5 gpurun_out/capture_spmm_v2.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_arxiv.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out
