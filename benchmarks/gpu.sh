#!/bin/bash
# One parametrised runner for everything this repo sends to a B200 box through gpurun (outputs land in gpurun_out/,
# the summaries worth keeping are copied to profiles/ by hand).  Sub-commands can be chained:
#
#   gpurun --timeout 1200 -- 'bash benchmarks/gpu.sh tests smoke bench:ogbn-arxiv'
#
#   tests[:expr]            pytest -m gpu (optionally -k expr)
#   smoke                   __graft_entry__.smoke()
#   bench:<workload>[:P]    bench.py --workload <workload> [--precision P]      (1 GPU)
#   mbench:<N>:<workload>   torchrun --nproc-per-node N bench.py --gpus N --workload <workload>
#   ref:<workload>          bench.py --impl reference
#   launches:<workload>     ncu launch list (gpu__time_duration) of a short bench run
#   pge[:quick][:ncu]       fused PGE kernels: accuracy / timing (benchmarks/pge_fused_check.py), optional ncu --set full
#   spmm[:ncu]              standalone SpMM sweep (benchmarks/spmm_sweep.py)
#   mtests:<N>              the NCCL tests (they spawn one worker per GPU themselves) on a box with N GPUs
TAG=${GS_TAG:-r2}
mkdir -p gpurun_out
for cmd in "$@"; do
  IFS=: read -r what a b c <<< "$cmd"
  case $what in
    tests)
      log=gpurun_out/${TAG}_pytest_gpu.log
      ( time timeout 1500 python -m pytest tests -q -m gpu -x ${a:+-k "$a"} ) > $log 2>&1; echo "pytest rc=$?" >> $log
      tail -8 $log | cut -c1-300 ;;
    smoke)
      log=gpurun_out/${TAG}_smoke.log
      ( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $log 2>&1; echo "smoke rc=$?" >> $log
      tail -4 $log | cut -c1-300 ;;
    bench)
      log=gpurun_out/${TAG}_bench_${a}${b:+_p$b}.log
      ( time timeout 900 python bench.py --workload $a ${b:+--precision $b} --steps ${GS_STEPS:-10} --warmup 3 ${GS_BENCH_FLAGS} ) > $log 2>&1
      echo "bench rc=$?" >> $log; grep '^{"metric' $log | tail -1 > gpurun_out/${TAG}_bench_${a}${b:+_p$b}.json
      python - "$log" <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith('{"metric'):
        d = json.loads(ln)
        print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"], "phases", d.get("phases"))
        print("roofline", {k: v for k, v in (d.get("roofline") or {}).items() if k in ("kernel", "bound", "frac", "avg_ms")})
PY
      tail -3 $log | cut -c1-200 ;;
    mbench)
      log=gpurun_out/${TAG}_bench${a}_${b}.log
      ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $a --master-addr 127.0.0.1 --master-port 29514 \
          bench.py --gpus $a --workload $b --steps ${GS_STEPS:-5} --warmup 3 ${GS_BENCH_FLAGS} ) > $log 2>&1
      echo "bench rc=$?" >> $log; grep '^{"metric' $log | tail -1 > gpurun_out/${TAG}_bench${a}_${b}.json
      grep '^{"metric' $log | tail -1 | cut -c1-400; tail -3 $log | cut -c1-200 ;;
    ref)
      log=gpurun_out/${TAG}_bench_ref_${a}.log
      ( time timeout 900 python bench.py --impl reference --workload $a --steps 3 --warmup 1 ) > $log 2>&1; tail -3 $log | cut -c1-600 ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
        --log-file gpurun_out/${TAG}_launches_${a}.csv python bench.py --workload $a --steps 1 --warmup 1 --no-cpu-baseline --no-extra \
        > gpurun_out/${TAG}_bench_under_ncu_${a}.log 2>&1
      echo "launch list rc=$? ($(wc -l < gpurun_out/${TAG}_launches_${a}.csv) lines)" ;;
    pge)
      bash benchmarks/run_pge_fused_check.sh ${a:-full} $b | tail -12 | cut -c1-900 ;;
    spmm)
      ( time timeout 1200 python benchmarks/spmm_sweep.py --out gpurun_out/${TAG}_spmm_sweep.json ${GS_SPMM_FLAGS} ) > gpurun_out/${TAG}_spmm_sweep.log 2>&1
      tail -30 gpurun_out/${TAG}_spmm_sweep.log | cut -c1-250 ;;
    mtests)
      # the NCCL tests spawn their own one-process-per-GPU workers (torch.multiprocessing): plain pytest on a box with
      # >= 2 GPUs.  (Under torchrun every rank would spawn its own group on the same GPUs and the groups deadlock.)
      log=gpurun_out/${TAG}_pytest_nccl_${a}.log
      ( time timeout 420 python -m pytest tests -q -m gpu -k "nccl or sharded or row_partition" -x ) > $log 2>&1; echo "rc=$?" >> $log
      tail -8 $log | cut -c1-300 ;;
    *) echo "unknown sub-command $cmd" ;;
  esac
done
