"""Workload for `ncu --set full`: the three tall-and-skinny products of the real side at the ogbn-arxiv shape
(n2 = 152,064 sampled rows over 40 classes, d = 128, h = 256, C = 40), precision 1, fused epilogues.

    ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 3 \
        -o gpurun_out/prof_real_gemm -f python profiles/capture_real_gemm.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphslim_b200.ops import CudaOps  # noqa: E402

K = CudaOps("cuda:0", precision=1)
n2, d, h, C = 152064, 128, 256, 40
Xg = torch.randn(n2, d, device="cuda")
W1 = torch.randn(d, h, device="cuda") * 0.1
b1 = torch.randn(h, device="cuda") * 0.1
W2 = torch.randn(h, C, device="cuda") * 0.1
b2 = torch.randn(C, device="cuda") * 0.1
dU = torch.randn(n2, C, device="cuda")


def step():
    H1 = K.gemm(Xg, W1, bias=b1, relu=True)          # 152k x 128 x 256, bias + ReLU epilogue
    U = K.gemm(H1, W2, bias=b2)                      # 152k x 256 x 40
    dA1 = K.gemm(dU, W2, tb=True, mask=H1)           # 152k x 40 x 256, masked epilogue
    return H1, U, dA1


step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record()
H1 = K.gemm(Xg, W1, bias=b1, relu=True)
ev[1].record()
U = K.gemm(H1, W2, bias=b2)
ev[2].record()
dA1 = K.gemm(dU, W2, tb=True, mask=H1)
ev[3].record()
torch.cuda.synchronize()
print("H1 %.1f us, U %.1f us, dA1 %.1f us (pack kernels included)" % tuple(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(3)))
