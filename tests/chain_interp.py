"""Plain-PyTorch interpreter of a recorded chain (graphslim_b200/chain.py) -- the reference csrc/chain.cu is tested
against on the GPU, and the check of the recorder itself on CPU."""
import torch

from graphslim_b200 import chain as C


def interpret(ops):
    """Executes ChainRecorder.ops (descriptor dicts holding the operand tensors) in order, in place."""
    for f, _, _ in ops:
        k = f["kind"]
        if k == C.GEMM:
            a = f["A"].T if f["ta"] else f["A"]
            b = f["B"].T if f["tb"] else f["B"]
            v = f["alpha"] * (a.double() @ b.double()).float()
            if f["beta"] != 0.0:
                v = v + f["beta"] * f["C"]
            if f.get("bias") is not None:
                v = v + f["bias"]
            if f["relu"]:
                v = torch.relu(v)
            if f.get("mask") is not None:
                v = v * (f["mask"] > 0)
            f["C"].copy_(v)
        elif k == C.SOFTMAX_RESIDUAL:
            Z = f["A"]
            ls = torch.log_softmax(Z, dim=1)
            S = ls.exp()
            Y = torch.zeros_like(S)
            Y[torch.arange(Z.shape[0], device=Z.device), f["B"].long()] = 1.0
            sc = f["bias"][:, None] if f.get("bias") is not None else 1.0
            f["C"].copy_(S)
            f["p5"].copy_((S - Y) * sc)
        elif k == C.COLSUM:
            f["C"].copy_(f["A"].double().sum(0, keepdim=True).float())
        elif k == C.ADAM_TABLE:
            p, g, m, v, table = f["C"], f["A"], f["p5"], f["p6"], f["B"]
            t = int(f["p7"].item())
            step_size, bc2_sqrt = float(table[t, 0]), float(table[t, 1])
            m.add_(f["alpha"] * (g - m))
            v.mul_(f["beta"]).add_(f["f0"] * g * g)
            p.sub_(step_size * (m / (v.sqrt() / bc2_sqrt + f["f1"])))
        elif k == C.COUNTER_ADD:
            f["C"].add_(f["M"])
        elif k == C.FILL:
            f["C"].fill_(f["alpha"])
        else:
            raise AssertionError(f"unknown op kind {k}")
