"""Workload for `ncu --set full` on the CSR SpMM (gs_spmm_csr_f32): power-law graphs of the BASELINE config-5 sweep.

    ncu --set full --clock-control none --import-source on -k regex:spmm_wide_kernel -o gpurun_out/prof_spmm \
        python profiles/capture_spmm.py

One launch per case (ncu flushes caches before every replay, so each capture is cold-cache like the L2-flushed sweep):
  0  ogbn-arxiv shape      n=169,343  nnz~2.5M  F=128   (X = 87 MB, fits L2)
  1  nnz 1e7, avg deg 8    n=1.25M             F=128   (X = 640 MB, DRAM-gather regime)
  2  nnz 1e7, avg deg 64   n=156k              F=256   (X = 160 MB)
  3  Reddit-like           n=203k  nnz 1e8     F=602   (X = 489 MB, 492 gathers per row: L2-bandwidth regime)
Prints the CUDA-event time of each (not a bench value when run under ncu).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
from spmm_sweep import powerlaw_csr  # noqa: E402
from graphslim_b200.graph_utils import build_row_chunks, chunks_to_device  # noqa: E402
from graphslim_b200.ops import Csr, CudaOps  # noqa: E402

dev = torch.device("cuda:0")
K = CudaOps(dev)
cases = [(169343, 2_500_000, 128), (1_250_000, 10**7, 128), (156_250, 10**7, 256), (203_252, 10**8, 602)]
if len(sys.argv) > 1:
    cases = [cases[int(a)] for a in sys.argv[1:]]
for n, nnz_t, F in cases:
    rowptr, col, val = powerlaw_csr(n, nnz_t, seed=7, dev=dev)
    chunks = chunks_to_device(build_row_chunks(rowptr.cpu().numpy(), 64), dev)
    csr = Csr(rowptr, col, val, n, n, chunks)
    ld = (F + 7) // 8 * 8
    X = torch.zeros(n, ld, device=dev)
    X[:, :F] = torch.randn(n, F, device=dev)
    Y = torch.zeros(n, ld, device=dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    K.spmm(csr, X, out=Y)
    b.record()
    torch.cuda.synchronize()
    nnz = col.numel()
    alg = 4 * (n + 1) + 8 * nnz + 8 * F * n
    gather = 4 * (n + 1) + 8 * nnz + 4 * F * nnz + 4 * F * n
    print(f"n={n} nnz={nnz} F={F} ms={a.elapsed_time(b):.3f} alg_bytes={alg} gather_bytes={gather}", flush=True)
    del csr, X, Y, rowptr, col, val
    torch.cuda.empty_cache()
