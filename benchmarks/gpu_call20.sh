#!/bin/bash
mkdir -p gpurun_out
( time timeout 240 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_arxiv_v9.log 2>&1
grep '^{"metric' gpurun_out/bench_arxiv_v9.log | tail -1 | cut -c1-200
