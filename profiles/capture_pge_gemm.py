"""Workload for `ncu --set full`: the PGE layer-2 products at the ogbn-arxiv shape (N'=909, h=256), precision 1.

    ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 3 -o gpurun_out/prof_pge_gemm \
        python profiles/capture_pge_gemm.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphslim_b200.ops import CudaOps  # noqa: E402

K = CudaOps("cuda:0", precision=1)
n, h = 909, 256
H1 = torch.randn(n * n, h, device="cuda").clamp_(min=0)
dY = torch.randn(n * n, h, device="cuda")
W = torch.randn(h, h, device="cuda") * 0.06
for _ in range(2):
    K.gemm(H1, W, tb=True)          # forward  Y2 = H1 W2^T
K.gemm(H1, W, tb=True)
K.gemm(dY, W)                       # backward dH1 = dY2 W2
K.gemm(dY, H1, ta=True)             # backward dW2 = dY2^T H1
torch.cuda.synchronize()
