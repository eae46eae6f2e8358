"""Stand-alone check of the tcgen05 GEMM (gs_gemm_f32 with precision 1 / 2) against a float64 product.

Run as a subprocess with a timeout by tests/test_gemm_tc_gpu.py so that a deadlocked mbarrier pipeline shows up as a
failed test instead of a hung test session:   python -m tests.tc_check [quick|full|bench]
"""
import sys
import time

import torch


def run(mode):
    from graphslim_b200.ops import CudaOps
    K = CudaOps("cuda:0")
    gen = torch.Generator().manual_seed(0)
    shapes = [
        # (M, N, K, ta, tb)
        (128, 256, 64, False, True),
        (128, 256, 256, False, True),
        (256, 256, 256, False, True),
        (300, 256, 256, False, True),
        (4900, 128, 128, False, True),
        (1000, 40, 256, False, False),
        (909, 1600, 909, False, False),
        (909, 1600, 909, True, False),
        (256, 256, 20000, True, False),
        (128, 300, 100, False, False),
        (70, 49, 1433, True, True),
        (513, 257, 130, False, True),
    ]
    if mode == "quick":
        shapes = shapes[:4]
    worst = {1: 0.0, 2: 0.0}
    for (M, N, Kd, ta, tb) in shapes:
        A = torch.randn((Kd, M) if ta else (M, Kd), generator=gen)
        B = torch.randn((N, Kd) if tb else (Kd, N), generator=gen)
        ref = ((A.T if ta else A).double() @ (B.T if tb else B).double())
        scale = ref.abs().max().item()
        for prec in (1, 2):
            C0 = torch.randn(M, N, generator=gen)
            out = C0.clone().cuda()
            K.gemm(A.cuda(), B.cuda(), ta=ta, tb=tb, out=out, alpha=0.5, beta=2.0, precision=prec)
            torch.cuda.synchronize()
            want = 0.5 * ref + 2.0 * C0.double()
            err = (out.cpu().double() - want).abs().max().item() / scale
            worst[prec] = max(worst[prec], err)
            print(f"M={M} N={N} K={Kd} ta={ta} tb={tb} precision={prec}: max err / max|C| = {err:.3e}", flush=True)
            tol = 2e-5 if prec == 1 else 2e-2
            if not err < tol:
                print("FAIL", flush=True)
                return 1
    # grouped K-segmented product (per-class weight gradients) with 64-aligned, partly empty segments
    seg = torch.tensor([0, 640, 640, 6400, 6464, 13120], dtype=torch.int32)
    ob = torch.tensor([3, 0, 1, 4, 2], dtype=torch.int32)
    for (M, N) in ((128, 256), (256, 40), (602, 41)):
        A = torch.randn(13120, M, generator=gen)
        B = torch.randn(13120, N, generator=gen)
        ref = torch.zeros(M, 6 * N, dtype=torch.float64)
        sl = seg.tolist()
        for g, o in enumerate(ob.tolist()):
            ref[:, o * N:(o + 1) * N] = A[sl[g]:sl[g + 1]].double().T @ B[sl[g]:sl[g + 1]].double()
        for prec in (1, 2):
            out = K.gemm_grouped_tn(A.cuda(), B.cuda(), seg.cuda(), ob.cuda(), 6, aligned=True, precision=prec)
            torch.cuda.synchronize()
            err = (out.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
            print(f"grouped M={M} N={N} precision={prec}: max err / max|C| = {err:.3e}", flush=True)
            if not err < (2e-5 if prec == 1 else 2e-2):
                print("FAIL", flush=True)
                return 1
    print(f"worst: 3xBF16 {worst[1]:.3e}, BF16 {worst[2]:.3e}")
    if mode == "bench":
        n, h = 909, 256
        Y = torch.randn(n * n, h, device="cuda")
        H = torch.randn(n * n, h, device="cuda")
        Wm = torch.randn(h, h, device="cuda")
        for name, fn in (("dW = dY^T H   (256x256, K=N'^2, TN)", lambda pr: K.gemm(Y, H, ta=True, precision=pr)),
                         ("dX = dY W     (N'^2x256x256, NN)", lambda pr: K.gemm(Y, Wm, precision=pr))):
            for prec in (1,):
                fn(prec)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    fn(prec)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 5
                print(f"{name} precision={prec}: {ms:.3f} ms, {2.0*n*n*h*h/ms/1e9:.1f} TFLOP/s", flush=True)
        del Y, H
        A = torch.randn(n * n, h, device="cuda")
        W = torch.randn(h, h, device="cuda")
        out = torch.empty(n * n, h, device="cuda")
        for prec in (0, 1, 2):
            for _ in range(2):
                K.gemm(A, W, tb=True, out=out, precision=prec)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                K.gemm(A, W, tb=True, out=out, precision=prec)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            print(f"PGE layer-2 shape {n*n}x{h}x{h} precision={prec}: {ms:.3f} ms, "
                  f"{2.0*n*n*h*h/ms/1e9:.1f} TFLOP/s (algorithmic)", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(run(sys.argv[1] if len(sys.argv) > 1 else "full"))
