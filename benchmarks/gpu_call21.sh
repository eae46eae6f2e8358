#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu_v4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v4.log; tail -6 gpurun_out/pytest_gpu_v4.log | cut -c1-300
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_v2.log 2>&1; tail -4 gpurun_out/smoke_v2.log | cut -c1-200
