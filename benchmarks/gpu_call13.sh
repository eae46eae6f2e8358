#!/bin/bash
# per-shape SIMT vs tcgen05 timings of the small workloads (drives the gemm_tc_covers threshold) + flaky-test recheck
mkdir -p gpurun_out
for w in cora flickr reddit; do
  timeout 300 python benchmarks/gemm_shapes.py --workload $w --out gpurun_out/gemm_shapes_$w.json > gpurun_out/gemm_shapes_$w.log 2>&1
  echo "$w: $(tail -1 gpurun_out/gemm_shapes_$w.log)"
done
( time timeout 600 python -m pytest tests/test_gcond_gpu.py -q -k graph ) > gpurun_out/pytest_graph2.log 2>&1
tail -3 gpurun_out/pytest_graph2.log
