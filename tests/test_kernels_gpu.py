"""Every CUDA kernel (called through the C ABI via graphslim_b200.ops.CudaOps) against its plain-PyTorch fp32
reference in tests/emu_ops.py.  fp32 kernels: rtol 1e-4 unless stated; integer/structure outputs exact."""
import numpy as np
import pytest
import torch

from graphslim_b200.ops import Csr
from tests.emu_ops import EmuOps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from graphslim_b200.ops import CudaOps
    return CudaOps("cuda:0")


@pytest.fixture(scope="module")
def E():
    return EmuOps("cpu")


def close(a, b, rtol=1e-4, atol=None):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    if atol is None:
        atol = rtol * (b.abs().max().item() + 1e-30)
    assert a.shape == b.shape, (a.shape, b.shape)
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def rand_csr(n_rows, n_cols, avg_deg, gen, powerlaw=False, dev="cpu"):
    if powerlaw:
        deg = np.minimum((gen.pareto(1.2, n_rows) * avg_deg / 3 + 1).astype(np.int64), n_cols)
    else:
        deg = gen.integers(0, 2 * avg_deg + 1, n_rows)
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(deg)
    col = np.concatenate([np.sort(gen.choice(n_cols, d, replace=False)) for d in deg]) if deg.sum() else \
        np.zeros(0, np.int64)
    val = gen.standard_normal(col.size).astype(np.float32)
    mk = lambda a, dt: torch.from_numpy(a.astype(dt)).to(dev)
    return Csr(mk(rowptr, np.int32), mk(col, np.int32), mk(val, np.float32), n_rows, n_cols)


def to_dev(csr, dev):
    ch = None if csr.chunks is None else tuple(c.to(dev) for c in csr.chunks)
    return Csr(csr.rowptr.to(dev), csr.col.to(dev), csr.val.to(dev), csr.n_rows, csr.n_cols, ch)


# ------------------------------------------------------------------------------------------ SpMM
@pytest.mark.parametrize("F", [1, 7, 40, 41, 128, 256, 500, 602, 1433])
@pytest.mark.parametrize("padded", [False, True])
def test_spmm_widths(K, E, F, padded):
    gen = np.random.default_rng(F)
    csr = rand_csr(777, 913, 9, gen)
    ld = (F + 7) // 8 * 8 if padded else F
    Xh = torch.zeros(913, ld)
    Xh[:, :F] = torch.from_numpy(gen.standard_normal((913, F)).astype(np.float32))
    ref = E.spmm(csr, Xh[:, :F])
    got = K.spmm(to_dev(csr, "cuda"), Xh.cuda()[:, :F])
    close(got, ref)


def test_spmm_empty_rows_and_accumulate(K, E):
    gen = np.random.default_rng(0)
    csr = rand_csr(300, 300, 3, gen)
    X = torch.from_numpy(gen.standard_normal((300, 64)).astype(np.float32))
    base = torch.from_numpy(gen.standard_normal((300, 64)).astype(np.float32))
    ref = base + E.spmm(csr, X)
    out = base.clone().cuda()
    K.spmm(to_dev(csr, "cuda"), X.cuda(), out=out, accumulate=True)
    close(out, ref)
    # zero rows
    empty = Csr(torch.zeros(1, dtype=torch.int32), torch.zeros(0, dtype=torch.int32), torch.zeros(0), 0, 300)
    assert K.spmm(to_dev(empty, "cuda"), X.cuda()).shape == (0, 64)


def test_spmm_long_rows_chunked(K, E):
    from graphslim_b200.graph_utils import build_row_chunks, chunks_to_device
    gen = np.random.default_rng(5)
    csr = rand_csr(2000, 5000, 8, gen, powerlaw=True)
    X = torch.from_numpy(gen.standard_normal((5000, 128)).astype(np.float32))
    ref = E.spmm(csr, X)
    d = to_dev(csr, "cuda")
    d.chunks = chunks_to_device(build_row_chunks(csr.rowptr.numpy(), 64), "cuda")
    assert d.chunks is not None
    got = K.spmm(d, X.cuda())
    close(got, ref)


def _spmm_seq_fp32(csr, X):
    """Row by row, non-zeros in CSR order, one fused multiply-add per element: the order gs_spmm_csr_f32 promises.
    float64 fma emulation is exact for fp32 products, so rounding once per step reproduces fmaf."""
    rp, col, val = csr.rowptr.numpy(), csr.col.numpy(), csr.val.numpy().astype(np.float64)
    Xd = X.numpy().astype(np.float64)
    out = np.zeros((csr.n_rows, X.shape[1]), dtype=np.float32)
    for r in range(csr.n_rows):
        acc = np.zeros(X.shape[1], dtype=np.float32)
        for e in range(rp[r], rp[r + 1]):
            acc = (val[e] * Xd[col[e]] + acc.astype(np.float64)).astype(np.float32)
        out[r] = acc
    return torch.from_numpy(out)


@pytest.mark.parametrize("n_rows", [1, 7, 40000])
def test_spmm_row_order_bit_exact(K, n_rows):
    """Rows of 0..200 non-zeros (multi-slice rows, empty rows, slices of exactly 32): the result equals the sequential
    CSR-order fp32 fused-multiply-add accumulation bit for bit."""
    gen = np.random.default_rng(n_rows)
    n_cols = 3000
    deg = gen.choice([0, 1, 2, 5, 31, 32, 33, 64, 65, 200], n_rows, p=[.1, .2, .2, .2, .05, .05, .05, .05, .05, .05])
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(deg)
    col = gen.integers(0, n_cols, rowptr[-1])
    val = gen.standard_normal(col.size).astype(np.float32)
    csr = Csr(torch.from_numpy(rowptr.astype(np.int32)), torch.from_numpy(col.astype(np.int32)), torch.from_numpy(val),
              n_rows, n_cols)
    X = torch.from_numpy(gen.standard_normal((n_cols, 128)).astype(np.float32))
    got = K.spmm(to_dev(csr, "cuda"), X.cuda()).cpu()
    sub = slice(0, min(n_rows, 600))
    ref = _spmm_seq_fp32(Csr(csr.rowptr[: sub.stop + 1], csr.col, csr.val, sub.stop, n_cols), X)
    assert torch.equal(got[sub], ref)
    tail = Csr(csr.rowptr[-201:] if n_rows > 200 else csr.rowptr, csr.col, csr.val, min(n_rows, 200), n_cols)
    assert torch.equal(got[-tail.n_rows:], _spmm_seq_fp32(tail, X))


@pytest.mark.parametrize("F", [128, 130, 256, 384, 608, 1024, 1436])
def test_spmm_tunings_bit_identical(K, F):
    """Every gs_spmm_set_tuning setting (kernel generation, gathers in flight, rows per warp, cache hints) promises the
    same bits: CSR-order fmaf accumulation for whole rows; rows sliced for atomics are compared with a tolerance."""
    from graphslim_b200.graph_utils import build_row_chunks, chunks_to_device
    gen = np.random.default_rng(F)
    n_rows, n_cols = 5000, 4000
    deg = gen.choice([0, 1, 2, 5, 31, 32, 33, 64, 65, 97, 300], n_rows, p=[.1, .2, .2, .2, .05, .05, .05, .05, .04, .03, .03])
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(deg)
    col = gen.integers(0, n_cols, rowptr[-1])
    val = gen.standard_normal(col.size).astype(np.float32)
    csr = to_dev(Csr(torch.from_numpy(rowptr.astype(np.int32)), torch.from_numpy(col.astype(np.int32)),
                     torch.from_numpy(val), n_rows, n_cols), "cuda")
    X = torch.from_numpy(gen.standard_normal((n_cols, F)).astype(np.float32)).cuda()
    base0 = torch.from_numpy(gen.standard_normal((n_rows, F)).astype(np.float32)).cuda()
    tunings = [(1, 4, 0, 0, 8, 8), (1, 0, 0, 0, 0, 0), (1, 8, 0, 3, 4, 8), (1, 4, 0, 4, 2, 4), (1, 8, 0, 7, 8, 2),
               (1, 4, 0, 5, 4, 1),
               (2, 0, 0, 3, 8, 8), (2, 2, 0, 3, 8, 8), (2, 4, 1, 0, 8, 8), (2, 8, 3, 1, 8, 8), (2, 0, 16, 2, 8, 8),
               (2, 4, 32, 3, 8, 8)]
    try:
        outs = []
        for t in tunings:
            K.spmm_set_tuning(*t)
            plain = K.spmm(csr, X)
            acc = K.spmm(csr, X, out=base0.clone(), accumulate=True)
            csr.chunks = chunks_to_device(build_row_chunks(rowptr, 64), "cuda")
            sliced = K.spmm(csr, X)
            csr.chunks = None
            outs.append((plain, acc, sliced))
        for plain, acc, sliced in outs[1:]:
            assert torch.equal(plain, outs[0][0])
            assert torch.equal(acc, outs[0][1])
            short = torch.from_numpy(deg <= 64).cuda()
            assert torch.equal(sliced[short], outs[0][0][short])
            close(sliced.cpu(), outs[0][0].cpu())
        sub = Csr(csr.rowptr[:301].cpu(), csr.col.cpu(), csr.val.cpu(), 300, n_cols)
        assert torch.equal(outs[0][0][:300].cpu(), _spmm_seq_fp32(sub, X.cpu()))
    finally:
        K.spmm_auto_tuning()


def test_spmm_transpose_and_scatter_agree(K, E):
    gen = np.random.default_rng(7)
    csr = rand_csr(400, 650, 6, gen)
    dY = torch.from_numpy(gen.standard_normal((400, 40)).astype(np.float32))
    ref = E.spmm_scatter(csr, dY, torch.zeros(650, 40))
    got = K.spmm_scatter(to_dev(csr, "cuda"), dY.cuda(), torch.zeros(650, 40, device="cuda"))
    close(got, ref)


def test_spmm_l2_column_tiling_bit_identical(K):
    """gs_spmm_csr_tiled_f32 (X swept in L2-resident column slices), with and without long-row work items, for slice widths
    that do and do not divide F.  Slices of >= 128 floats run the wide kernel (every output element accumulates its
    non-zeros in CSR order, as untiled); narrower slices -- including the remainder slice -- run the split-warp kernel
    whose lane groups take alternate non-zeros, so the results agree to fp32 reassociation, and exactly on the columns of
    the full-width wide slices."""
    from graphslim_b200.graph_utils import build_row_chunks, chunks_to_device
    from graphslim_b200.ops import Csr
    gen = torch.Generator().manual_seed(5)
    n, F = 3000, 602
    deg = torch.randint(1, 40, (n,), generator=gen)
    deg[7] = 900                                            # a hub row
    rowptr = torch.cat([torch.zeros(1, dtype=torch.int64), deg.cumsum(0)]).to(torch.int32)
    nnz = int(rowptr[-1])
    col = torch.randint(0, n, (nnz,), generator=gen).to(torch.int32)
    val = torch.rand(nnz, generator=gen)
    Xp = torch.zeros(n, 608)
    Xp[:, :F] = torch.randn(n, F, generator=gen)
    X = Xp.cuda()[:, :F]
    chunks = chunks_to_device(build_row_chunks(rowptr.numpy(), 64), "cuda")
    for ch in (None, chunks):
        csr = Csr(rowptr.cuda(), col.cuda(), val.cuda(), n, n, ch)
        ref = K.spmm(csr, X, tile_cols=0)
        for tile in (32, 64, 96, 128, 256):
            got = K.spmm(csr, X, tile_cols=tile)
            close(got, ref, rtol=1e-5)
            if tile >= 128 and ch is None:
                full = (F // tile) * tile
                assert torch.equal(got[:, :full], ref[:, :full])
    # the automatic choice: nothing to tile for a matrix that fits the L2
    assert K.spmm_tile_cols(csr, X) == 0


def test_gather_rows(K):
    gen = np.random.default_rng(1)
    X = torch.from_numpy(gen.standard_normal((1000, 602)).astype(np.float32)).cuda()
    idx = torch.from_numpy(gen.integers(0, 1000, 5000).astype(np.int32)).cuda()
    assert torch.equal(K.gather_rows(X, idx), X[idx.long()])


def test_csr_gcn_norm_bit_exact(K, E):
    gen = np.random.default_rng(2)
    csr = rand_csr(3000, 3000, 12, gen)
    a = torch.ones_like(csr.val)
    r = torch.from_numpy(np.power(gen.integers(1, 50, 3000).astype(np.float64), -0.5))
    ref = E.csr_gcn_norm(csr.rowptr, csr.col, a, r)
    got = K.csr_gcn_norm(csr.rowptr.cuda(), csr.col.cuda(), a.cuda(), r.cuda())
    assert torch.equal(got.cpu(), ref)


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K_", [(1, 1, 1), (70, 7, 1433), (257, 129, 65), (909, 1600, 40), (128, 128, 4096),
                                     (1, 300, 909), (300, 1, 17), (640, 512, 256)])
def test_gemm_shapes(K, E, ta, tb, M, N, K_):
    gen = torch.Generator().manual_seed(M * 31 + N * 7 + K_)
    A = torch.randn((K_, M) if ta else (M, K_), generator=gen)
    B = torch.randn((N, K_) if tb else (K_, N), generator=gen)
    ref = E.gemm(A, B, ta, tb)
    got = K.gemm(A.cuda(), B.cuda(), ta, tb)
    close(got, ref, rtol=2e-5)


def test_gemm_beta_alpha_views_splitk(K, E):
    gen = torch.Generator().manual_seed(3)
    W = torch.randn(64, 200, generator=gen)
    X = torch.randn(50, 100, generator=gen)
    C0 = torch.randn(50, 64, generator=gen)
    ref = 0.5 * (X @ W[:, 100:].T) + 2.0 * C0
    out = C0.clone().cuda()
    K.gemm(X.cuda(), W.cuda()[:, 100:], tb=True, out=out, alpha=0.5, beta=2.0)
    close(out, ref, rtol=2e-5)
    # strided output views
    dW = torch.empty(64, 200, device="cuda")
    P = torch.randn(50, 64, generator=gen)
    K.gemm(P.cuda(), X.cuda(), ta=True, out=dW[:, :100])
    K.gemm(P.cuda(), X.cuda(), ta=True, out=dW[:, 100:], alpha=-1.0)
    close(dW, torch.cat([P.T @ X, -(P.T @ X)], 1), rtol=2e-5)
    # split-K path (small output, huge K) with beta accumulate
    A = torch.randn(30000, 96, generator=gen)
    B = torch.randn(30000, 80, generator=gen)
    C1 = torch.randn(96, 80, generator=gen)
    out = C1.clone().cuda()
    K.gemm(A.cuda(), B.cuda(), ta=True, out=out, beta=1.0)
    close(out, C1 + (A.double().T @ B.double()).float(), rtol=5e-5)


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("M,N,K_,ta,tb", [
    (446, 7, 256, False, False),       # H1 W2 at the Flickr shape: warp-per-row, float4 lanes
    (446, 7, 446, False, False),       # A_hat M2: lda = 446, scalar lanes
    (7297, 7, 256, False, True),       # tall A, B given transposed
    (70, 7, 1433, False, False),       # K spans three shared-memory chunks, unaligned rows
    (33, 16, 520, False, False),       # N = 16 (two outputs per lane), chunk tail of 8
    (5, 9, 4, False, False),           # K <= 16 wins over N <= 16
    (256, 7, 446, True, False),        # H1^T dM2: lanes over the columns of A, warps over K
    (1433, 3, 70, True, False),
    (500, 16, 1000, True, False),
    (7296, 256, 7, False, True),       # dM2 W2^T: short dot products
    (446, 49, 1, False, False),        # bias outer product
    (3122, 257, 7, False, False),
    (300, 1800, 16, True, True),
])
def test_gemm_skinny(M, N, K_, ta, tb, precision, E):
    """Products with one dimension <= 16 (csrc/gemm.cu: skinny kernels) incl. alpha/beta and the fused epilogue; they are
    exact-fp32 at every gemm_precision."""
    from graphslim_b200.ops import CudaOps
    K = CudaOps("cuda:0", precision=precision)
    gen = torch.Generator().manual_seed(M * 13 + N * 5 + K_)
    A = torch.randn((K_, M) if ta else (M, K_), generator=gen)
    B = torch.randn((N, K_) if tb else (K_, N), generator=gen)
    ref = E.gemm(A, B, ta, tb)
    close(K.gemm(A.cuda(), B.cuda(), ta, tb), ref, rtol=2e-5)
    C0 = torch.randn(M, N, generator=gen)
    out = C0.clone().cuda()
    K.gemm(A.cuda(), B.cuda(), ta, tb, out=out, alpha=0.5, beta=2.0)
    close(out, 0.5 * ref + 2.0 * C0, rtol=2e-5)
    bias, mask = torch.randn(N, generator=gen), torch.randn(M, N, generator=gen)
    got = K.gemm(A.cuda(), B.cuda(), ta, tb, bias=bias.cuda(), relu=True, mask=mask.cuda())
    close(got, torch.relu(ref + bias) * (mask > 0), rtol=2e-5)
    # strided operands and output (views into wider buffers)
    Aw = torch.randn(A.shape[0], A.shape[1] + 5, generator=gen)
    Cw = torch.zeros(M, N + 3).cuda()
    K.gemm(Aw.cuda()[:, 2:2 + A.shape[1]], B.cuda(), ta, tb, out=Cw[:, 1:1 + N])
    close(Cw[:, 1:1 + N], E.gemm(Aw[:, 2:2 + A.shape[1]].contiguous(), B, ta, tb), rtol=2e-5)
    assert float(Cw[:, 0].abs().max()) == 0.0 and float(Cw[:, 1 + N:].abs().max()) == 0.0


@pytest.mark.parametrize("M,N", [(500, 7), (1433, 7), (256, 16), (77, 9)])
def test_gemm_grouped_tn_skinny(K, E, M, N):
    gen = torch.Generator().manual_seed(M + N)
    seg = torch.tensor([0, 64, 64, 1100, 1101, 4000], dtype=torch.int32)
    ob = torch.tensor([3, 0, 1, 5, 4], dtype=torch.int32)
    A = torch.randn(4000, M, generator=gen)
    B = torch.randn(4000, N, generator=gen)
    ref = E.gemm_grouped_tn(A, B, seg, ob, 7)
    got = K.gemm_grouped_tn(A.cuda(), B.cuda(), seg.cuda(), ob.cuda(), 7)
    close(got, ref, rtol=2e-5)
    again = K.gemm_grouped_tn(A.cuda(), B.cuda(), seg.cuda(), ob.cuda(), 7)
    assert torch.equal(got, again)                          # no atomics: bit-reproducible


def test_gemm_grouped_tn(K, E):
    gen = torch.Generator().manual_seed(4)
    seg = torch.tensor([0, 10, 10, 300, 1000], dtype=torch.int32)
    ob = torch.tensor([3, 0, 1, 4], dtype=torch.int32)
    A = torch.randn(1000, 77, generator=gen)
    B = torch.randn(1000, 9, generator=gen)
    ref = E.gemm_grouped_tn(A, B, seg, ob, 6)
    got = K.gemm_grouped_tn(A.cuda(), B.cuda(), seg.cuda(), ob.cuda(), 6)
    close(got, ref, rtol=2e-5)


def test_segment_colsum(K, E):
    gen = torch.Generator().manual_seed(41)
    seg = torch.tensor([0, 64, 64, 700, 1000], dtype=torch.int32)
    ob = torch.tensor([2, 0, 4, 1], dtype=torch.int32)
    for cols in (1, 40, 256, 257):
        X = torch.randn(1000, cols, generator=gen)
        close(K.segment_colsum(X.cuda(), seg.cuda(), ob.cuda(), 5), E.segment_colsum(X, seg, ob, 5), rtol=1e-5)


# ------------------------------------------------------------------------------------------ glue kernels
def test_bias_relu_mask(K, E):
    gen = torch.Generator().manual_seed(5)
    Z = torch.randn(301, 257, generator=gen)
    b = torch.randn(257, generator=gen)
    close(K.bias_act(Z.clone().cuda(), b.cuda(), True), E.bias_act(Z.clone(), b, True), rtol=1e-6)
    close(K.bias_act(Z.clone().cuda(), b.cuda(), False), E.bias_act(Z.clone(), b, False), rtol=1e-6)
    H = torch.randn(301, 33, generator=gen)
    D = torch.randn(301, 5 * 33, generator=gen)
    close(K.relu_mask(D.clone().cuda(), H.cuda(), groups=5), E.relu_mask(D.clone(), H, groups=5), rtol=0, atol=0)


@pytest.mark.parametrize("C", [2, 7, 40, 41, 100])
def test_softmax_family(K, E, C):
    gen = torch.Generator().manual_seed(C)
    rows = 513
    Z = torch.randn(rows, C, generator=gen) * 3
    lab = torch.randint(0, C, (rows,), generator=gen, dtype=torch.int32)
    sc = torch.rand(rows, generator=gen)
    S, R, nll = K.softmax_residual(Z.cuda(), lab.cuda(), sc.cuda(), want_nll=True)
    Sr, Rr, nr = E.softmax_residual(Z, lab, sc, want_nll=True)
    close(S, Sr, rtol=1e-5)
    close(R, Rr, rtol=1e-5)
    close(nll, nr, rtol=1e-5)
    blk = torch.randint(0, 6, (rows,), generator=gen, dtype=torch.int32)
    Ex = K.expand_class_blocks(Rr.cuda(), blk.cuda(), 6)
    assert torch.equal(Ex.cpu(), E.expand_class_blocks(Rr, blk, 6))
    Zf = torch.randn(rows, 6 * C, generator=gen)
    assert torch.equal(K.pick_class_blocks(Zf.cuda(), blk.cuda(), 6).cpu(), E.pick_class_blocks(Zf, blk, 6))
    Q = torch.randn(rows, C, generator=gen)
    close(K.softmax_jvp(Sr.cuda(), Q.cuda(), sc.cuda()), E.softmax_jvp(Sr, Q, sc), rtol=1e-5)


@pytest.mark.parametrize("metric", ["ours", "mse", "cos"])
def test_match(K, E, metric):
    gen = torch.Generator().manual_seed(11)
    nc, widths, rows, is_bias = 5, [16, 16, 5, 5], [48, 1, 16, 1], [False, True, False, True]
    gs = [torch.randn(r, nc * w, generator=gen) for r, w in zip(rows, widths)]
    gr = [torch.randn(r, nc * w, generator=gen) for r, w in zip(rows, widths)]
    gs[0][:, 3] = 0.0                                   # zero-norm column: torch's norm backward gives 0 there
    coeff = torch.rand(nc, generator=gen)
    le, lk = torch.zeros(1), torch.zeros(1, device="cuda")
    Ge = E.match(gs, gr, widths, is_bias, coeff, metric, le)
    Gk = K.match([g.cuda() for g in gs], [g.cuda() for g in gr], widths, is_bias, coeff.cuda(), metric, lk)
    close(lk, le, rtol=1e-5)
    for a, b in zip(Gk, Ge):
        close(a, b, rtol=1e-4)


def test_dense_gcn_norm(K, E):
    gen = torch.Generator().manual_seed(12)
    A = torch.rand(153, 153, generator=gen)
    A = (A + A.T) / 2
    A.fill_diagonal_(0)
    Ah, r = K.dense_gcn_norm(A.cuda())
    Ahr, rr = E.dense_gcn_norm(A)
    close(Ah, Ahr, rtol=1e-5)
    close(r, rr, rtol=1e-5)
    dAh = torch.randn(153, 153, generator=gen)
    close(K.dense_gcn_norm_bwd(dAh.cuda(), Ahr.cuda(), rr.cuda()), E.dense_gcn_norm_bwd(dAh, Ahr, rr), rtol=1e-4)


def test_adam_matches_torch(K):
    gen = torch.Generator().manual_seed(13)
    p0 = torch.randn(1000, generator=gen)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p_ref], lr=0.01)
    p, m, v = p0.clone().cuda(), torch.zeros(1000).cuda(), torch.zeros(1000).cuda()
    for t in range(1, 6):
        g = torch.randn(1000, generator=gen)
        p_ref.grad = g.clone()
        opt.step()
        K.adam_step(p, g.cuda(), m, v, t, 0.01)
    close(p, p_ref.detach(), rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------------------ PGE kernels
@pytest.mark.parametrize("n,h,cuts", [(37, 128, [0, 12, 37]), (64, 256, [0, 16, 32, 48, 64]), (70, 128, [0, 23, 46, 70])])
def test_pge_row_sharded_pieces_match_unsharded(K, n, h, cuts):
    """The per-rank halves of the row-sharded PGE (expand rows, partial / combined BatchNorm statistics, BN1 backward
    pass / final) reproduce the unsharded kernels when the slices are combined the way the collectives combine them."""
    gen = torch.Generator().manual_seed(n * h)
    dev = "cuda"
    Pa = torch.randn(n, h, generator=gen).to(dev)
    Pb = (torch.randn(n, h, generator=gen) + 0.5).to(dev)
    gamma = (torch.rand(h, generator=gen) + 0.5).to(dev)
    beta = (torch.randn(h, generator=gen) * 0.1).to(dev)
    off = torch.tensor([0, n * n], dtype=torch.int64, device=dev)
    offs = [torch.tensor([0, (b - a) * n], dtype=torch.int64, device=dev) for a, b in zip(cuts[:-1], cuts[1:])]
    mean1, rstd1, cm1 = K.pge_l1_stats_closed(Pa, Pb)
    # layer-1 expansion: bit-identical rows
    H1 = K.pge_l1_expand(Pa, Pb, off, mean1, rstd1, gamma, beta)
    H1s = torch.cat([K.pge_l1_expand_rows(Pa, Pb[a:b], o, mean1, rstd1, gamma, beta)
                     for (a, b), o in zip(zip(cuts[:-1], cuts[1:]), offs)])
    assert torch.equal(H1, H1s)
    # column statistics: shifted partial sums merged in double
    Y = (torch.randn(n * n, h, generator=gen) * 3 + 1).to(dev)
    mean, rstd = K.col_stats_chunked(Y, off)
    parts = torch.stack([torch.cat([K.col_stats_partial(Y[a * n:b * n], o), Y[a * n].double()])
                         for (a, b), o in zip(zip(cuts[:-1], cuts[1:]), offs)])
    counts = torch.tensor([(b - a) * n for a, b in zip(cuts[:-1], cuts[1:])], dtype=torch.int64, device=dev)
    mean_s, rstd_s = K.col_stats_combine(parts, counts)
    close(mean_s, mean, rtol=1e-6)
    close(rstd_s, rstd, rtol=1e-5)
    # BN1 backward: per-slice linear reductions, summed like the all-reduce does
    dH1 = torch.randn(n * n, h, generator=gen).to(dev)
    ref = K.pge_bn1_bwd_closed(dH1, Pa, Pb, mean1, rstd1, gamma, beta, cm1)
    work = None
    for a, b in zip(cuts[:-1], cuts[1:]):
        w = K.pge_bn1_bwd_pass_rows(dH1[a * n:b * n], Pa, Pb, a, b - a, mean1, rstd1, gamma, beta)
        if work is None:
            work = w
        else:
            work[:2 * h] += w[:2 * h]
            work[2 * h:].view(torch.float32).add_(w[2 * h:].view(torch.float32))
    got = K.pge_bn1_bwd_final(Pa, Pb, rstd1, gamma, cm1, work)
    for g_, r_ in zip(got, ref):
        close(g_, r_, rtol=1e-4)


@pytest.mark.parametrize("n,h,nchunk", [(70, 128, 1), (40, 128, 5), (153, 256, 1), (97, 128, 5)])
def test_pge_kernels(K, E, n, h, nchunk):
    gen = torch.Generator().manual_seed(n + h + nchunk)
    total = n * n
    sizes = [total // nchunk + (1 if i < total % nchunk else 0) for i in range(nchunk)]
    off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64)
    Pa = torch.randn(n, h, generator=gen)
    Pb = torch.randn(n, h, generator=gen) + 0.5
    gamma = torch.rand(h, generator=gen) + 0.5
    beta = torch.randn(h, generator=gen) * 0.1
    c = lambda t: t.cuda()
    m1, r1 = K.pge_l1_stats(c(Pa), c(Pb), c(off))
    m1r, r1r = E.pge_l1_stats(Pa, Pb, off)
    close(m1, m1r, rtol=1e-5)
    close(r1, r1r, rtol=1e-5)
    H1 = K.pge_l1_expand(c(Pa), c(Pb), c(off), c(m1r), c(r1r), c(gamma), c(beta))
    H1r = E.pge_l1_expand(Pa, Pb, off, m1r, r1r, gamma, beta)
    close(H1, H1r, rtol=1e-5)
    Y2 = torch.randn(total, h, generator=gen) * 2 + 1
    m2, r2 = K.col_stats_chunked(c(Y2), c(off))
    m2r, r2r = E.col_stats_chunked(Y2, off)
    close(m2, m2r, rtol=1e-5)
    close(r2, r2r, rtol=1e-5)
    w3 = torch.randn(h, generator=gen) * 0.1
    b3 = torch.randn(1, generator=gen)
    Ek = K.pge_l3(c(Y2), c(off), c(m2r), c(r2r), c(gamma), c(beta), c(w3), c(b3))
    Er = E.pge_l3(Y2, off, m2r, r2r, gamma, beta, w3, b3)
    close(Ek, Er, rtol=1e-4)
    A = K.pge_symm_sigmoid(c(Er), n)
    Ar = E.pge_symm_sigmoid(Er, n)
    close(A, Ar, rtol=1e-5)
    dA = torch.randn(n, n, generator=gen)
    dE = K.pge_symm_sigmoid_bwd(c(dA), c(Ar))
    dEr = E.pge_symm_sigmoid_bwd(dA, Ar)
    close(dE, dEr, rtol=1e-5)
    s1, s2, dw3, db3 = K.pge_l3_bwd_stats(c(Y2), c(dEr), c(off), c(m2r), c(r2r), c(gamma), c(beta), c(w3))
    s1r, s2r, dw3r, db3r = E.pge_l3_bwd_stats(Y2, dEr, off, m2r, r2r, gamma, beta, w3)
    for a, b in ((s1, s1r), (s2, s2r), (dw3, dw3r), (db3, db3r)):
        close(a, b, rtol=1e-4)
    dY2 = K.pge_bn2_bwd_apply(c(Y2), c(dEr), c(off), c(m2r), c(r2r), c(gamma), c(beta), c(w3), c(s1r), c(s2r))
    close(dY2, E.pge_bn2_bwd_apply(Y2, dEr, off, m2r, r2r, gamma, beta, w3, s1r, s2r), rtol=1e-4)
    dH1 = torch.randn(total, h, generator=gen)
    t1, t2 = K.pge_bn1_bwd_stats(c(dH1), c(Pa), c(Pb), c(off), c(m1r), c(r1r), c(gamma), c(beta))
    t1r, t2r = E.pge_bn1_bwd_stats(dH1, Pa, Pb, off, m1r, r1r, gamma, beta)
    close(t1, t1r, rtol=1e-4)
    close(t2, t2r, rtol=1e-4)
    dPa, dPb = K.pge_bn1_bwd_reduce(c(dH1), c(Pa), c(Pb), c(off), c(m1r), c(r1r), c(gamma), c(beta), c(t1r), c(t2r))
    dPar, dPbr = E.pge_bn1_bwd_reduce(dH1, Pa, Pb, off, m1r, r1r, gamma, beta, t1r, t2r)
    close(dPa, dPar, rtol=2e-4)
    close(dPb, dPbr, rtol=2e-4)


def test_cpu_tensors_are_rejected(K):
    with pytest.raises((TypeError, ValueError, RuntimeError)):
        K.gemm(torch.randn(4, 4), torch.randn(4, 4))


@pytest.mark.parametrize("n,h", [(23, 128), (61, 256), (40, 64)])
def test_pge_unchunked_fast_paths(K, E, n, h):
    """Closed-form BN1 statistics, one-pass BN1 backward and the register-resident layer-3 kernel against the generic
    per-chunk kernels' reference."""
    gen = torch.Generator().manual_seed(7 * n + h)
    total = n * n
    off = torch.tensor([0, total], dtype=torch.int64)
    Pa = torch.randn(n, h, generator=gen) * 1.5 + 0.3
    Pb = torch.randn(n, h, generator=gen) - 0.7
    gamma = torch.rand(h, generator=gen) + 0.5
    beta = torch.randn(h, generator=gen) * 0.1
    c = lambda t: t.cuda()
    m1r, r1r = E.pge_l1_stats(Pa, Pb, off)
    m1, r1, cm = K.pge_l1_stats_closed(c(Pa), c(Pb))
    close(m1, m1r, rtol=1e-5)
    close(r1, r1r, rtol=1e-5)
    close(cm, torch.stack([Pa.double().mean(0), Pb.double().mean(0)]).float(), rtol=1e-5)
    dH1 = torch.randn(total, h, generator=gen)
    t1r, t2r = E.pge_bn1_bwd_stats(dH1, Pa, Pb, off, m1r, r1r, gamma, beta)
    dPar, dPbr = E.pge_bn1_bwd_reduce(dH1, Pa, Pb, off, m1r, r1r, gamma, beta, t1r, t2r)
    dPa, dPb, dg, db = K.pge_bn1_bwd_closed(c(dH1), c(Pa), c(Pb), c(m1r), c(r1r), c(gamma), c(beta), cm)
    close(dPa, dPar, rtol=1e-4)
    close(dPb, dPbr, rtol=1e-4)
    close(dg, t2r.sum(0), rtol=1e-4)
    close(db, t1r.sum(0), rtol=1e-4)
    Y2 = torch.randn(total, h, generator=gen) * 2 + 1
    m2r, r2r = E.col_stats_chunked(Y2, off)
    w3 = torch.randn(h, generator=gen) * 0.1
    b3 = torch.randn(1, generator=gen)
    Ek = K.pge_l3(c(Y2), c(off), c(m2r), c(r2r), c(gamma), c(beta), c(w3), c(b3))
    close(Ek, E.pge_l3(Y2, off, m2r, r2r, gamma, beta, w3, b3), rtol=1e-4)
