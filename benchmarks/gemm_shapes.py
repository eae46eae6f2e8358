#!/usr/bin/env python
"""Times every distinct dense product of one GCond epoch at a workload shape with each gs_gemm_f32 precision
(0 = SIMT fp32, 1 = tcgen05 3xBF16 split), to drive the per-shape dispatch in graphslim_b200/ops.py.

    python benchmarks/gemm_shapes.py [--workload ogbn-arxiv] [--out gpurun_out/gemm_shapes.json]
"""
import argparse
import collections
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ogbn-arxiv")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gemm_shapes.json"))
    ns = ap.parse_args()
    from graphslim_b200.reduction import create_reducer
    raw, args, gdata = bench.make_problem(ns.workload, 0, epochs=1, gemm_precision=1, track_loss=False)
    data = gdata.TransAndInd(raw, args.dataset, args.pre_norm)
    bench.seed_everything(args.seed)
    agent = create_reducer(args.method, setting=args.setting, data=data, args=args)
    agent.setup(data)
    K = agent.K
    seen = collections.Counter()
    orig = K.gemm

    def spy(A, B, ta=False, tb=False, out=None, alpha=1.0, beta=0.0, precision=None, **epi):
        M, Kk = (A.shape[1], A.shape[0]) if ta else (A.shape[0], A.shape[1])
        N = B.shape[0] if tb else B.shape[1]
        seen[(int(ta), int(tb), M, N, Kk, float(beta) != 0.0)] += 1
        return orig(A, B, ta, tb, out, alpha, beta, precision, **epi)

    K.gemm = spy
    agent.run_epoch(0)
    torch.cuda.synchronize()
    K.gemm = orig
    rows = []
    for (ta, tb, M, N, Kk, acc), count in sorted(seen.items(), key=lambda kv: -kv[1]):
        A = torch.randn((Kk, M) if ta else (M, Kk), device="cuda")
        B = torch.randn((N, Kk) if tb else (Kk, N), device="cuda")
        C = torch.zeros(M, N, device="cuda")
        rec = dict(ta=ta, tb=tb, M=M, N=N, K=Kk, beta=acc, calls_per_epoch=count)
        for prec in (0, 1):
            ts = []
            for i in range(12):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                orig(A, B, bool(ta), bool(tb), out=C, beta=1.0 if acc else 0.0, precision=prec)
                b.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(a.elapsed_time(b) * 1e3)
            rec[f"us_p{prec}"] = round(statistics.median(ts), 1)
        rows.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs(os.path.dirname(ns.out), exist_ok=True)
    json.dump(rows, open(ns.out, "w"), indent=1)
    tot0 = sum(r["us_p0"] * r["calls_per_epoch"] for r in rows)
    tot1 = sum(r["us_p1"] * r["calls_per_epoch"] for r in rows)
    best = sum(min(r["us_p0"], r["us_p1"]) * r["calls_per_epoch"] for r in rows)
    print(f"per epoch: all-SIMT {tot0 / 1e3:.1f} ms, all-tcgen05 {tot1 / 1e3:.1f} ms, best-of {best / 1e3:.1f} ms")


if __name__ == "__main__":
    main()
