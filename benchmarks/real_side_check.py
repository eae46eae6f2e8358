"""Timing of the real-side hidden-layer backward at the arxiv shape: fused grouped kernels (csrc/grouped_tn.cu) against
the pack_b + generic tcgen05 path they replace.    python benchmarks/real_side_check.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphslim_b200.ops import CudaOps          # noqa: E402


def timeit(fn, iters=10):
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
    K = CudaOps("cuda", precision=1)
    G, per, d, h, C = 40, 3840, 128, 256, 40                  # 40 classes x ~3.8 K sampled rows (arxiv shape)
    R = G * per
    seg = torch.arange(0, R + 1, per, dtype=torch.int32, device="cuda")
    ids = torch.arange(G, dtype=torch.int32, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn(R, d, device="cuda", generator=gen)
    H1 = torch.relu(torch.randn(R, h, device="cuda", generator=gen))
    dU = torch.randn(R, C, device="cuda", generator=gen)
    W2 = torch.randn(h, C, device="cuda", generator=gen)
    out = {}
    out["ms_gW2_mn"] = timeit(lambda: K.gemm_grouped_tn(H1, dU, seg, ids, G, aligned=True))
    out["ms_gW1_gb1_fused"] = timeit(lambda: K.mlp_bwd_grouped(X, H1, dU, W2, seg, ids, G))
    K.grouped_mn = False

    def old():
        dA1 = K.gemm(dU, W2, tb=True, mask=H1)
        K.gemm_grouped_tn(X, dA1, seg, ids, G, aligned=True)
        K.segment_colsum(dA1, seg, ids, G)
    out["ms_gW2_old"] = timeit(lambda: K.gemm_grouped_tn(H1, dU, seg, ids, G, aligned=True))
    out["ms_gW1_gb1_old"] = timeit(old)
    print(json.dumps(out))


if __name__ == "__main__" and "--forward" not in sys.argv:
    main()


def forward_probe():
    """The two forward products of the real side at the arxiv shape, tensor-core (3xBF16) vs fp32 SIMT."""
    R, d, h, C = 40 * 3840, 128, 256, 40
    gen = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(R, d, device="cuda", generator=gen)
    W1 = torch.randn(d, h, device="cuda", generator=gen)
    b1 = torch.randn(h, device="cuda", generator=gen)
    W2 = torch.randn(h, C, device="cuda", generator=gen)
    b2 = torch.randn(C, device="cuda", generator=gen)
    out = {}
    for prec in (1, 0):
        K = CudaOps("cuda", precision=prec)
        H1 = K.gemm(X, W1, bias=b1, relu=True)
        out[f"ms_H1_p{prec}"] = timeit(lambda: K.gemm(X, W1, bias=b1, relu=True))
        out[f"ms_U_p{prec}"] = timeit(lambda: K.gemm(H1, W2, bias=b2))
    print(json.dumps(out))


if __name__ == "__main__" and "--forward" in sys.argv:
    forward_probe()
