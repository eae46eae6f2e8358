#!/bin/bash
# Every fused-PGE stage in its own process under a timeout (a hung mbarrier wait must not take the call down).
# usage: run_pge_fused_check.sh [quick|full] [ncu]
mode=${1:-full}
out=gpurun_out/pge_fused_check.log
: > $out
run() { echo "== $*" >> $out; timeout 180 python benchmarks/pge_fused_check.py "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
if [ "$mode" = full ]; then
  for stage in fwd dx_store dx_fused dw; do
    run --stage $stage --n 70 --h 128
    run --stage $stage --n 61 --h 256
  done
  run --stage all --n 23 --h 128 --precision 2
fi
run --stage all --n 153 --h 256
run --stage all --n 97 --h 256 --i-first 20 --n-i 41
run --stage all --n 909 --h 256 --time --no-ref
run --stage all --n 446 --h 256 --time --no-ref
if [ "$2" = ncu ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pge_l2_ -c 6 -f -o gpurun_out/r2_pge_fused \
    python benchmarks/pge_fused_check.py --stage all --n 909 --h 256 --no-ref >> $out 2>&1
fi
cat $out
