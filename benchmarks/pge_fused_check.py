"""Diagnostics for csrc/pge_fused.cu on a GPU: every fused kernel against the plain-PyTorch reference (tests/emu_ops.py)
and, with --time, against the unfused kernels it replaces.  One stage per process (--stage) so a hang in one kernel
cannot hide the others; run under `timeout`.

    python benchmarks/pge_fused_check.py --stage fwd --n 70 --h 128
    python benchmarks/pge_fused_check.py --stage all --n 909 --h 256 --time
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from graphslim_b200.ops import CudaOps          # noqa: E402
from tests.emu_ops import EmuOps                # noqa: E402


def relerr(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def make_inputs(n, h, seed):
    gen = torch.Generator().manual_seed(seed)
    Pa = torch.randn(n, h, generator=gen) * 1.3 + 0.2
    Pb = torch.randn(n, h, generator=gen) - 0.4
    gamma1 = torch.rand(h, generator=gen) + 0.5
    beta1 = torch.randn(h, generator=gen) * 0.2
    gamma2 = torch.rand(h, generator=gen) + 0.5
    beta2 = torch.randn(h, generator=gen) * 0.2
    W2 = torch.randn(h, h, generator=gen) / h ** 0.5
    w3 = torch.randn(h, generator=gen) * 0.1
    dE = torch.randn(n * n, generator=gen)
    return dict(Pa=Pa, Pb=Pb, gamma1=gamma1, beta1=beta1, gamma2=gamma2, beta2=beta2, W2=W2, w3=w3, dE=dE)


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default="all", choices=["fwd", "dx_store", "dx_fused", "dw", "all"])
    ap.add_argument("--n", type=int, default=70)
    ap.add_argument("--h", type=int, default=128)
    ap.add_argument("--i-first", type=int, default=0)
    ap.add_argument("--n-i", type=int, default=0)
    ap.add_argument("--precision", type=int, default=1)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--no-ref", action="store_true", help="skip the CPU reference (large shapes: timing only)")
    args = ap.parse_args()
    n, h = args.n, args.h
    i0 = args.i_first
    n_i = args.n_i or (n - i0)
    K = CudaOps("cuda", precision=args.precision)
    E = EmuOps("cpu")
    inp = make_inputs(n, h, 1000 * n + h)
    c = {k: v.cuda() for k, v in inp.items()}
    out = dict(n=n, h=h, i_first=i0, n_i=n_i, precision=args.precision)
    off = torch.tensor([0, n_i * n], dtype=torch.int64, device="cuda")
    mean1, rstd1, cm1 = K.pge_l1_stats_closed(c["Pa"], c["Pb"])
    bn1 = (mean1, rstd1, c["gamma1"], c["beta1"])
    bn1_cpu = tuple(t.cpu() for t in bn1)
    stages = ["fwd", "dx_store", "dx_fused", "dw"] if args.stage == "all" else [args.stage]
    count = float(n) * float(n)

    # reference forward (CPU fp32 -> the values the later stages consume are the CUDA ones, compared separately)
    if not args.no_ref:
        Y2_ref, stats_ref = E.pge_fused_l2_fwd(inp["Pa"], inp["Pb"], i0, n_i, *bn1_cpu, inp["W2"])
    Y2, stats = K.pge_fused_l2_fwd(c["Pa"], c["Pb"], i0, n_i, *bn1, c["W2"])
    torch.cuda.synchronize()
    if "fwd" in stages:
        if not args.no_ref:
            out["fwd_Y2"] = relerr(Y2, Y2_ref)
            out["fwd_stats_sum"] = relerr(stats[:h], stats_ref[:h])
            out["fwd_stats_sq"] = relerr(stats[h:], stats_ref[h:])
        # the unfused CUDA path on the same inputs
        H1 = K.pge_l1_expand_rows(c["Pa"], c["Pb"][i0:i0 + n_i], off, *bn1)
        Y2_old = K.gemm(H1, c["W2"], tb=True)
        out["fwd_vs_unfused"] = relerr(Y2, Y2_old)
        if args.time:
            out["ms_fwd_fused"] = timeit(lambda: K.pge_fused_l2_fwd(c["Pa"], c["Pb"], i0, n_i, *bn1, c["W2"]))

            def old():
                H = K.pge_l1_expand_rows(c["Pa"], c["Pb"][i0:i0 + n_i], off, *bn1)
                Y = K.gemm(H, c["W2"], tb=True)
                K.col_stats_chunked(Y, off)
            out["ms_fwd_unfused"] = timeit(old)
        del H1, Y2_old
    # the backward stages work on the SAME Y2 (the CUDA one) so their errors are their own
    mean2, rstd2 = K.pge_stats_finalize(stats, float(n_i) * n)       # statistics of the slice (enough for a kernel check)
    bn2 = (mean2, rstd2, c["gamma2"], c["beta2"])
    dE = c["dE"][i0 * n:(i0 + n_i) * n].contiguous()
    s1, s2, _, _ = K.pge_l3_bwd_stats(Y2, dE, off, mean2, rstd2, c["gamma2"], c["beta2"], c["w3"])
    torch.cuda.synchronize()
    cpu = lambda t: t.detach().cpu()
    bn2_cpu = tuple(cpu(t) for t in bn2)
    if not args.no_ref:
        Y2c, dEc, s1c, s2c = cpu(Y2), cpu(dE), cpu(s1), cpu(s2)
    if "dx_store" in stages:
        dH1 = K.pge_fused_l2_bwd_dx(c["Pa"], c["Pb"], i0, n_i, bn1, c["W2"], Y2, dE, bn2, c["w3"], s1, s2, count,
                                    store=True)
        torch.cuda.synchronize()
        if not args.no_ref:
            ref = E.pge_fused_l2_bwd_dx(inp["Pa"], inp["Pb"], i0, n_i, bn1_cpu, inp["W2"], Y2c, dEc, bn2_cpu, inp["w3"],
                                        s1c, s2c, count, store=True)
            out["dx_store_dH1"] = relerr(dH1, ref)
        del dH1
    if "dx_fused" in stages:
        work = K.pge_bn1_work(n, h)
        K.pge_fused_l2_bwd_dx(c["Pa"], c["Pb"], i0, n_i, bn1, c["W2"], Y2, dE, bn2, c["w3"], s1, s2, count, work=work)
        K.pge_bn1_tsum(c["Pa"], c["Pb"], cm1, rstd1, work)
        torch.cuda.synchronize()
        if not args.no_ref:
            wr = E.pge_bn1_work(n, h)
            E.pge_fused_l2_bwd_dx(inp["Pa"], inp["Pb"], i0, n_i, bn1_cpu, inp["W2"], Y2c, dEc, bn2_cpu, inp["w3"], s1c,
                                  s2c, count, work=wr)
            E.pge_bn1_tsum(inp["Pa"], inp["Pb"], cpu(cm1), cpu(rstd1), wr)
            fl, flr = cpu(work)[2 * h:].view(torch.float32), wr[2 * h:].view(torch.float32)
            out["dx_fused_Ga"] = relerr(fl[:n * h], flr[:n * h])
            out["dx_fused_Gb"] = relerr(fl[n * h:], flr[n * h:])
            out["dx_fused_t1"] = relerr(cpu(work)[:h], wr[:h])
            out["dx_fused_t2"] = relerr(cpu(work)[h:2 * h], wr[h:2 * h])
        if args.time:
            def new():
                w = K.pge_bn1_work(n, h)
                K.pge_fused_l2_bwd_dx(c["Pa"], c["Pb"], i0, n_i, bn1, c["W2"], Y2, dE, bn2, c["w3"], s1, s2, count,
                                      work=w)
            out["ms_dx_fused"] = timeit(new)
            out["ms_dx_store"] = timeit(lambda: K.pge_fused_l2_bwd_dx(c["Pa"], c["Pb"], i0, n_i, bn1, c["W2"], Y2, dE,
                                                                      bn2, c["w3"], s1, s2, count, store=True))

            def old():
                dY2 = K.pge_bn2_bwd_apply(Y2, dE, off, mean2, rstd2, c["gamma2"], c["beta2"], c["w3"], s1, s2)
                d = K.gemm(dY2, c["W2"])
                K.pge_bn1_bwd_pass_rows(d, c["Pa"], c["Pb"], i0, n_i, *bn1)
            out["ms_dx_unfused"] = timeit(old)
    if "dw" in stages:
        dW2 = K.pge_fused_l2_bwd_dw(c["Pa"], c["Pb"], i0, n_i, bn1, Y2, dE, bn2, c["w3"], s1, s2, count)
        torch.cuda.synchronize()
        if not args.no_ref:
            ref = E.pge_fused_l2_bwd_dw(inp["Pa"], inp["Pb"], i0, n_i, bn1_cpu, Y2c, dEc, bn2_cpu, inp["w3"], s1c, s2c,
                                        count)
            out["dw_dW2"] = relerr(dW2, ref)
        if args.time:
            out["ms_dw_fused"] = timeit(lambda: K.pge_fused_l2_bwd_dw(c["Pa"], c["Pb"], i0, n_i, bn1, Y2, dE, bn2,
                                                                      c["w3"], s1, s2, count))

            def old():
                dY2 = K.pge_bn2_bwd_apply(Y2, dE, off, mean2, rstd2, c["gamma2"], c["beta2"], c["w3"], s1, s2)
                H = K.pge_l1_expand_rows(c["Pa"], c["Pb"][i0:i0 + n_i], off, *bn1)
                K.gemm(dY2, H, ta=True)
            out["ms_dw_unfused_incl_inputs"] = timeit(old)
    print(json.dumps(out))


if __name__ == "__main__":
    t0 = time.time()
    main()
