#!/usr/bin/env python
"""BASELINE.json configs[4]: standalone A_hat @ X propagation sweep (power-law graphs, fp32 values, int32 indices).

    python benchmarks/spmm_sweep.py [--quick] [--out profiles/r1_spmm_sweep.json]

For every (nnz, avg degree, F) it reports the CUDA-event time of gs_spmm_csr_f32 (L2 flushed between iterations),
  ALG    = 4(N+1) + 8 nnz + 4 F N + 4 F N      bytes / time   (compulsory traffic; <= 100 % of the HBM roofline)
  GATHER = 4(N+1) + 8 nnz + 4 F nnz + 4 F N    bytes / time   (one feature-row gather per non-zero: what an untiled
                                                                 row-wise kernel moves through L2)
against the measured copy peak (MEASURED_PEAKS.json), and the same product through torch.sparse CSR mm on the same
GPU -- a STAND-IN for the reference's torch_sparse path, which is not installable here.
Graphs are generated on the GPU (test infrastructure: torch ops are fine here), symmetric, GCN-normalised.
"""
import argparse
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphslim_b200.graph_utils import build_row_chunks, chunks_to_device  # noqa: E402
from graphslim_b200.ops import Csr, CudaOps  # noqa: E402


def powerlaw_csr(n, nnz_target, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    w = torch.arange(1, n + 1, device=dev, dtype=torch.float64) ** (-1.0 / 1.3)
    w = w[torch.randperm(n, device=dev, generator=g)]
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    need = nnz_target // 2
    keys = torch.empty(0, dtype=torch.int64, device=dev)
    while keys.numel() < need:
        m = int((need - keys.numel()) * 1.4) + 64
        u = torch.searchsorted(cdf, torch.rand(m, device=dev, dtype=torch.float64, generator=g)).clamp_(max=n - 1)
        v = torch.searchsorted(cdf, torch.rand(m, device=dev, dtype=torch.float64, generator=g)).clamp_(max=n - 1)
        keep = u != v
        lo, hi = torch.minimum(u[keep], v[keep]), torch.maximum(u[keep], v[keep])
        keys = torch.unique(torch.cat([keys, lo * n + hi]))
    keys = keys[torch.randperm(keys.numel(), device=dev, generator=g)[:need]]
    lo, hi = keys // n, keys % n
    diag = torch.arange(n, device=dev)
    row = torch.cat([lo, hi, diag])
    col = torch.cat([hi, lo, diag])
    order = torch.argsort(row * n + col)
    row, col = row[order], col[order]
    deg = torch.bincount(row, minlength=n).double()
    r = deg.pow(-0.5)
    val = (r[row] * r[col]).float()
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    return rowptr.to(torch.int32), col.to(torch.int32), val


VARIANTS = {   # gs_spmm_set_tuning(impl, unr, group, flags, wpb, max_nv); 0 = auto
    "v1_w8_u4": (1, 4, 0, 0, 8, 8),
    "v1_w4_u8": (1, 8, 0, 0, 4, 8),
    "v1_w2_u4": (1, 4, 0, 0, 2, 8),
    "v1_w2_u8": (1, 8, 0, 0, 2, 8),
    "v1_auto": (1, 0, 0, 0, 0, 0),
    "v1_nv2_u8": (1, 8, 0, 0, 0, 2),
    "v1_nv2_u8_w4": (1, 8, 0, 0, 4, 2),
    "v1_nv2_u8_w2": (1, 8, 0, 0, 2, 2),
    "v1_nv3_u8": (1, 8, 0, 0, 0, 4),
    "v2": (2, 0, 0, 3, 0, 0),
}


def time_it(fn, flush, iters=8):
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--variants", action="store_true", help="also time every tuning of the wide kernel (VARIANTS)")
    ap.add_argument("--long-row", type=int, default=64, help="rows with more non-zeros are sliced (atomics)")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r1_spmm_sweep.json"))
    ns = ap.parse_args()
    dev = torch.device("cuda:0")
    K = CudaOps(dev)
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk_path))["hbm_gbs"] if os.path.exists(pk_path) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    grid = [(10**5, 8), (10**6, 8), (10**6, 64), (10**7, 8), (10**7, 64), (10**8, 64), (10**8, 492)]
    feats = [128, 256, 500, 602]
    if ns.quick:
        grid, feats = [(10**6, 8), (10**7, 8), (10**7, 64), (10**8, 492)], [128, 602]
    rows = []
    for nnz_t, avg in grid:
        n = max(1000, nnz_t // avg)
        rowptr, col, val = powerlaw_csr(n, nnz_t, seed=nnz_t % 97 + avg, dev=dev)
        nnz = col.numel()
        chunks = chunks_to_device(build_row_chunks(rowptr.cpu().numpy(), ns.long_row), dev)
        csr = Csr(rowptr, col, val, n, n, chunks)
        max_deg = int((rowptr[1:] - rowptr[:-1]).max())
        tcsr = torch.sparse_csr_tensor(rowptr.long(), col.long(), val, size=(n, n))
        for F in feats:
            ld = (F + 7) // 8 * 8
            Xp = torch.zeros(n, ld, device=dev)
            Xp[:, :F] = torch.randn(n, F, device=dev)
            X = Xp[:, :F]
            outp = torch.zeros(n, ld, device=dev)
            out = outp[:, :F]
            # rows are padded to 32 B with zeros, so the kernel runs at the padded width (float4 for any F)
            variants = {}
            if ns.variants:
                # (impl, unr, rows per warp, cache-hint flags) of gs_spmm_set_tuning; every variant must give the same bits
                base = None
                for name, tune in VARIANTS.items():
                    K.spmm_set_tuning(*tune)
                    outp.zero_()
                    variants[name] = time_it(lambda: K.spmm(csr, Xp, out=outp), flush, iters=5)
                    if base is None:
                        base = outp.clone()
                    elif not torch.equal(base, outp):
                        variants[name + "_MISMATCH"] = float((base - outp).abs().max())
                del base
                K.spmm_auto_tuning()
            ms = time_it(lambda: K.spmm(csr, Xp, out=outp), flush)
            ref = torch.sparse.mm(tcsr, X.contiguous())
            err = float((out - ref).abs().max() / ref.abs().max())
            Xc = X.contiguous()
            ms_t = time_it(lambda: torch.sparse.mm(tcsr, Xc), flush, iters=4)
            alg = 4 * (n + 1) + 8 * nnz + 4 * F * n + 4 * F * n
            gather = 4 * (n + 1) + 8 * nnz + 4 * F * nnz + 4 * F * n
            rec = dict(nnz=nnz, n=n, avg_deg=avg, max_deg=max_deg, F=F, ms=ms, alg_GBps=alg / ms / 1e6,
                       alg_frac_of_hbm=alg / ms / 1e6 / peak, gather_GBps=gather / ms / 1e6,
                       x_MB=4 * F * n / 1e6, torch_sparse_csr_mm_ms=ms_t, speedup_vs_torch_csr=ms_t / ms, max_rel_err=err)
            if variants:
                rec["variants_ms"] = variants
            rows.append(rec)
            print(json.dumps(rec), flush=True)
            del Xp, out, ref, Xc
        del tcsr, csr
        torch.cuda.empty_cache()
    json.dump(dict(hbm_peak_GBps=peak, long_row_nnz=ns.long_row, note="torch.sparse CSR mm is a stand-in for the reference's torch_sparse path",
                   rows=rows), open(ns.out, "w"), indent=1)


if __name__ == "__main__":
    main()
