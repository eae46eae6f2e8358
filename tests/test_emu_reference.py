"""The plain-PyTorch kernel references (tests/emu_ops.py) against torch autograd of the reference's own
formulas -- so that "CUDA kernel == emu" in the gpu tests means "CUDA kernel == reference maths".  CPU only."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import gcond_oracle as G
from tests.emu_ops import EmuOps

E = EmuOps("cpu")


@pytest.mark.parametrize("metric", ["ours", "mse", "cos"])
def test_match_gradient_equals_autograd_of_reference_formula(metric):
    gen = torch.Generator().manual_seed(0)
    nc, widths, rows, is_bias = 4, [6, 6, 3, 3], [10, 1, 6, 1], [False, True, False, True]
    gs = [torch.randn(r, nc * w, generator=gen, dtype=torch.float64).float().requires_grad_() for r, w in
          zip(rows, widths)]
    gr = [torch.randn(r, nc * w, generator=gen) for r, w in zip(rows, widths)]
    coeff = torch.rand(nc, generator=gen)
    total = torch.zeros(())
    for c in range(nc):
        syn = [(g[:, c * w:(c + 1) * w] if not b else g[0, c * w:(c + 1) * w]) for g, w, b in zip(gs, widths, is_bias)]
        real = [(g[:, c * w:(c + 1) * w] if not b else g[0, c * w:(c + 1) * w]) for g, w, b in zip(gr, widths, is_bias)]
        total = total + coeff[c] * G.match_loss(syn, real, metric)
    ref_grads = torch.autograd.grad(total, gs, allow_unused=True)
    loss = torch.zeros(1)
    got = E.match([g.detach() for g in gs], gr, widths, is_bias, coeff, metric, loss)
    torch.testing.assert_close(loss[0], total.detach(), rtol=1e-5, atol=1e-6)
    for a, b in zip(got, ref_grads):
        b = torch.zeros_like(a) if b is None else b
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-6)


def test_dense_norm_backward_equals_autograd():
    gen = torch.Generator().manual_seed(1)
    A = torch.rand(23, 23, generator=gen).requires_grad_()
    Ah = G.normalize_dense(A)
    dAh = torch.randn(23, 23, generator=gen)
    (ref,) = torch.autograd.grad((Ah * dAh).sum(), A)
    Ah2, r = E.dense_gcn_norm(A.detach())
    torch.testing.assert_close(Ah2, Ah.detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(E.dense_gcn_norm_bwd(dAh, Ah2, r), ref, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("dataset,rate,n", [("cora", 0.5, 13), ("reddit", 0.01, 11)])
def test_pge_forward_backward_equals_autograd(dataset, rate, n):
    """graphslim_b200.pge.PGE (on the emulated kernels) against the oracle's nn.Module restatement of PGE."""
    from types import SimpleNamespace
    from graphslim_b200.pge import PGE
    d = 9
    args = SimpleNamespace(dataset=dataset, reduction_rate=rate)
    torch.manual_seed(3)
    ref = G.PairwiseAdj(d, n, dataset, rate)
    torch.manual_seed(3)
    mine = PGE(E, d, n, args)
    for a, b in zip(mine.parameters(), ref.parameters()):
        assert torch.equal(a, b.detach())
    with torch.no_grad():                         # make BN affine non-trivial
        for bn in ref.bns:
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.2, 0.2)
    mine.gamma = [bn.weight.detach().clone() for bn in ref.bns]
    mine.beta = [bn.bias.detach().clone() for bn in ref.bns]
    x = torch.randn(n, d, requires_grad=True)
    A_ref = ref(x)
    dA = torch.randn(n, n)
    grads_ref = torch.autograd.grad((A_ref * dA).sum(), [x] + list(ref.parameters()))
    A = mine.forward(x.detach())
    torch.testing.assert_close(A, A_ref.detach(), rtol=1e-4, atol=1e-6)
    grads, dX = mine.backward(dA)
    torch.testing.assert_close(dX, grads_ref[0], rtol=1e-3, atol=1e-6)
    names = ["W1", "b1", "W2", "b2", "W3", "b3", "g1", "be1", "g2", "be2"]
    for name, a, b in zip(names, grads, grads_ref[1:]):
        if name in ("b1", "b2"):                 # cancelled by the following BatchNorm: exactly 0 here, fp noise there
            assert a.abs().max() == 0 and b.abs().max() < 1e-5
            continue
        torch.testing.assert_close(a.reshape(b.shape), b, rtol=1e-3, atol=1e-6 + 1e-4 * b.abs().max().item())
