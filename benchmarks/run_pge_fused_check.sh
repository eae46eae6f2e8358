#!/bin/bash
# Every fused-PGE stage in its own process under a timeout (a hung mbarrier wait must not take the call down).
out=gpurun_out/pge_fused_check.log
: > $out
run() { echo "== $*" >> $out; timeout 120 python benchmarks/pge_fused_check.py "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
for stage in fwd dx_store dx_fused dw; do
  run --stage $stage --n 70 --h 128
  run --stage $stage --n 61 --h 256
done
run --stage all --n 153 --h 256
run --stage all --n 23 --h 128 --precision 2
run --stage all --n 97 --h 256 --i-first 20 --n-i 41
run --stage all --n 909 --h 256 --time --no-ref
run --stage all --n 446 --h 256 --time --no-ref
cat $out
