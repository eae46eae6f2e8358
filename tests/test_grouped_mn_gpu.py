"""csrc/grouped_tn.cu (grouped TN products with MN-major, TMA-fed operands; the fused hidden-layer backward of the real
side) against the plain-PyTorch references, and the fused PGE kernels of csrc/pge_fused.cu against theirs."""
import numpy as np
import pytest
import torch

from tests.emu_ops import EmuOps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K1():
    from graphslim_b200.ops import CudaOps
    return CudaOps("cuda:0", precision=1)


@pytest.fixture(scope="module")
def E():
    return EmuOps("cpu")


def relerr(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def _segments(sizes):
    """64-aligned segment offsets (the sampler's padding) for true group sizes `sizes`; returns (seg, total, row mask)."""
    off, seg, live = 0, [0], []
    for s in sizes:
        pad = (s + 63) // 64 * 64
        live += [1.0] * s + [0.0] * (pad - s)
        off += pad
        seg.append(off)
    return torch.tensor(seg, dtype=torch.int32), off, torch.tensor(live)


@pytest.mark.parametrize("M,N", [(128, 256), (256, 40), (128, 128), (256, 256), (128, 8), (256, 64)])
def test_gemm_grouped_mn_matches_reference(K1, E, M, N):
    gen = torch.Generator().manual_seed(M * 7 + N)
    sizes = [200, 0, 64, 1000, 37, 513, 0, 90]                 # empty groups, one-stage groups, a long one
    seg, R, live = _segments(sizes)
    nblk = 10
    out_block = torch.tensor([3, 0, 9, 1, 4, 7, 2, 5], dtype=torch.int32)
    A = torch.randn(R, M, generator=gen)
    B = torch.randn(R, N, generator=gen) * live[:, None]        # padded rows contribute nothing
    ref = E.gemm_grouped_tn(A, B, seg, out_block, nblk)
    assert K1._mn_ok(A.cuda(), M, B.cuda(), N, M, N, 1)
    got = K1.gemm_grouped_tn(A.cuda(), B.cuda(), seg.cuda(), out_block.cuda(), nblk, aligned=True)
    assert relerr(got, ref) < 2e-5
    # strided operands (views into wider buffers) go through the same path
    Aw, Bw = torch.zeros(R, M + 8), torch.zeros(R, N + 4)
    Aw[:, :M], Bw[:, :N] = A, B
    got2 = K1.gemm_grouped_tn(Aw.cuda()[:, :M], Bw.cuda()[:, :N], seg.cuda(), out_block.cuda(), nblk, aligned=True)
    assert relerr(got2, ref) < 2e-5


@pytest.mark.parametrize("M,Cw", [(128, 40), (256, 8), (128, 64)])
def test_mlp_bwd_grouped_matches_reference(K1, E, M, Cw):
    gen = torch.Generator().manual_seed(M + Cw)
    sizes = [300, 64, 0, 1500, 65, 10]
    seg, R, live = _segments(sizes)
    nblk = 6
    out_block = torch.tensor([5, 1, 0, 2, 4, 3], dtype=torch.int32)
    h = 256
    X = torch.randn(R, M, generator=gen)
    H1 = torch.relu(torch.randn(R, h, generator=gen))
    dU = torch.randn(R, Cw, generator=gen) * live[:, None]
    W2 = torch.randn(h, Cw, generator=gen) * 0.2
    gW1r, gb1r = E.mlp_bwd_grouped(X, H1, dU, W2, seg, out_block, nblk)
    assert K1.mlp_bwd_grouped_supported(X.cuda(), H1.cuda(), dU.cuda(), True)
    gW1, gb1 = K1.mlp_bwd_grouped(X.cuda(), H1.cuda(), dU.cuda(), W2.cuda(), seg.cuda(), out_block.cuda(), nblk)
    assert relerr(gW1, gW1r) < 2e-5
    assert relerr(gb1, gb1r) < 2e-5


@pytest.mark.parametrize("n,h,i_first,n_i,precision", [(70, 128, 0, 70, 1), (61, 256, 0, 61, 1), (97, 256, 20, 41, 1),
                                                       (153, 256, 0, 153, 1), (23, 128, 0, 23, 2)])
def test_fused_pge_kernels_match_reference(E, n, h, i_first, n_i, precision):
    """pge_l2_fwd / pge_l2_bwd_dx (stored and fused-reduction epilogues) / pge_l2_bwd_dw on slices of the pair rows."""
    from graphslim_b200.ops import CudaOps
    K = CudaOps("cuda:0", precision=precision)
    tol = 3e-5 if precision == 1 else 1e-2
    gen = torch.Generator().manual_seed(1000 * n + h)
    Pa = torch.randn(n, h, generator=gen) * 1.3 + 0.2
    Pb = torch.randn(n, h, generator=gen) - 0.4
    g1, b1 = torch.rand(h, generator=gen) + 0.5, torch.randn(h, generator=gen) * 0.2
    g2, b2 = torch.rand(h, generator=gen) + 0.5, torch.randn(h, generator=gen) * 0.2
    W2 = torch.randn(h, h, generator=gen) / h ** 0.5
    w3 = torch.randn(h, generator=gen) * 0.1
    dE = torch.randn(n * n, generator=gen)[i_first * n:(i_first + n_i) * n].contiguous()
    c = lambda t: t.cuda()
    mean1, rstd1, cm1 = K.pge_l1_stats_closed(c(Pa), c(Pb))
    bn1 = (mean1, rstd1, c(g1), c(b1))
    bn1c = tuple(t.cpu() for t in bn1)
    Y2, stats = K.pge_fused_l2_fwd(c(Pa), c(Pb), i_first, n_i, *bn1, c(W2))
    Y2r, statsr = E.pge_fused_l2_fwd(Pa, Pb, i_first, n_i, *bn1c, W2)
    assert relerr(Y2, Y2r) < tol and relerr(stats, statsr) < tol
    count = float(n) * n
    mean2, rstd2 = K.pge_stats_finalize(stats, float(n_i) * n)
    bn2 = (mean2, rstd2, c(g2), c(b2))
    bn2c = tuple(t.cpu() for t in bn2)
    off = torch.tensor([0, n_i * n], dtype=torch.int64, device="cuda")
    s1, s2, _, _ = K.pge_l3_bwd_stats(Y2, c(dE), off, mean2, rstd2, c(g2), c(b2), c(w3))
    Y2c, s1c, s2c = Y2.cpu(), s1.cpu(), s2.cpu()
    dH1 = K.pge_fused_l2_bwd_dx(c(Pa), c(Pb), i_first, n_i, bn1, c(W2), Y2, c(dE), bn2, c(w3), s1, s2, count, store=True)
    dH1r = E.pge_fused_l2_bwd_dx(Pa, Pb, i_first, n_i, bn1c, W2, Y2c, dE, bn2c, w3, s1c, s2c, count, store=True)
    assert relerr(dH1, dH1r) < tol
    work = K.pge_bn1_work(n, h)
    K.pge_fused_l2_bwd_dx(c(Pa), c(Pb), i_first, n_i, bn1, c(W2), Y2, c(dE), bn2, c(w3), s1, s2, count, work=work)
    K.pge_bn1_tsum(c(Pa), c(Pb), cm1, rstd1, work)
    wr = E.pge_bn1_work(n, h)
    E.pge_fused_l2_bwd_dx(Pa, Pb, i_first, n_i, bn1c, W2, Y2c, dE, bn2c, w3, s1c, s2c, count, work=wr)
    E.pge_bn1_tsum(Pa, Pb, cm1.cpu(), rstd1.cpu(), wr)
    wk = work.cpu()
    assert relerr(wk[2 * h:].view(torch.float32), wr[2 * h:].view(torch.float32)) < tol     # Ga | Gb
    assert relerr(wk[:2 * h], wr[:2 * h]) < 2 * tol                                           # t1 | t2
    dW2 = K.pge_fused_l2_bwd_dw(c(Pa), c(Pb), i_first, n_i, bn1, Y2, c(dE), bn2, c(w3), s1, s2, count)
    dW2r = E.pge_fused_l2_bwd_dw(Pa, Pb, i_first, n_i, bn1c, Y2c, dE, bn2c, w3, s1c, s2c, count)
    assert relerr(dW2, dW2r) < tol
