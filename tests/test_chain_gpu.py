"""csrc/chain.cu (one persistent cooperative kernel walking a recorded operation list) against the plain-PyTorch
interpretation of the same list, and the inner loop run through it against the step-by-step path."""
import pytest
import torch

from graphslim_b200 import chain as C
from tests import helpers
from tests.chain_interp import interpret

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from graphslim_b200.ops import CudaOps
    return CudaOps("cuda:0", precision=1)


def _clone_ops(rec):
    """Deep copy of the recorded operand tensors on the CPU (same aliasing), for the reference interpretation."""
    memo = {}

    def cp(t):
        if t is None or not torch.is_tensor(t):
            return t
        base = t._base if t._base is not None else t
        key = base.data_ptr()
        if key not in memo:
            memo[key] = base.detach().cpu().clone()
        b = memo[key]
        return b.as_strided(t.shape, t.stride(), t.storage_offset() - base.storage_offset() + b.storage_offset())

    return [({k: cp(v) for k, v in f.items()}, r, w) for f, r, w in rec.ops], memo


@pytest.mark.parametrize("M,N,K_,ta,tb", [(909, 256, 128, False, False), (909, 40, 909, True, False),
                                           (256, 40, 909, True, False), (909, 256, 40, False, True),
                                           (70, 7, 1433, False, False), (33, 65, 31, True, True), (1, 1, 1, False, False),
                                           (446, 500, 446, False, True)])
def test_chain_gemm_epilogues(K, M, N, K_, ta, tb):
    gen = torch.Generator().manual_seed(M + 3 * N + 7 * K_)
    dev = "cuda:0"
    A = torch.randn((K_, M) if ta else (M, K_), generator=gen).to(dev)
    B = torch.randn((N, K_) if tb else (K_, N), generator=gen).to(dev)
    bias, mask = torch.randn(N, generator=gen).to(dev), torch.randn(M, N, generator=gen).to(dev)
    wide = torch.randn(M, N + 5, generator=gen).to(dev)
    rec = C.ChainRecorder(K)
    c0 = rec.gemm(A, B, ta, tb)
    c1 = rec.gemm(A, B, ta, tb, bias=bias, relu=True, mask=mask)
    c2 = rec.gemm(A, B, ta, tb, out=wide[:, 2:2 + N], alpha=0.5, beta=2.0)        # strided, accumulating
    s = rec.colsum(c1)
    ref_ops, memo = _clone_ops(rec)
    interpret(ref_ops)
    prog = rec.program()
    prog.run()
    torch.cuda.synchronize()
    for got, (f, _, _) in zip((c0, c1, c2, s), ref_ops):
        ref = f["C"]
        scale = float(ref.abs().max()) + 1e-30
        assert float((got.cpu() - ref).abs().max()) <= 2e-5 * scale
    assert torch.equal(wide[:, :2].cpu(), memo[wide.data_ptr()][:, :2])             # columns outside the view untouched
    prog.run()                                                                       # c2 accumulates again; others same
    torch.cuda.synchronize()
    assert float((c0.cpu() - ref_ops[0][0]["C"]).abs().max()) <= 2e-5 * (float(c0.abs().max()) + 1e-30)


def test_chain_softmax_adam_counter(K):
    gen = torch.Generator().manual_seed(11)
    dev = "cuda:0"
    Z = torch.randn(301, 41, generator=gen).to(dev)
    lab = torch.randint(0, 41, (301,), generator=gen).int().to(dev)
    sc = torch.rand(301, generator=gen).to(dev)
    p, g = torch.randn(1000, 33, generator=gen).to(dev), torch.randn(1000, 33, generator=gen).to(dev)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    table = K.adam_table(5, 0.01)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    rec = C.ChainRecorder(K)
    S, R = rec.softmax_residual(Z, lab, sc)
    z = rec.zeros(7, 9)
    rec.adam_step_table(p, g, m, v, table, step)
    rec.counter_add(step, 1)
    ref_ops, _ = _clone_ops(rec)
    prog = rec.program()
    for _ in range(3):
        interpret(ref_ops)
        prog.run()
    torch.cuda.synchronize()
    assert int(step.item()) == 3 and float(z.abs().max()) == 0.0
    torch.testing.assert_close(S.cpu(), ref_ops[0][0]["C"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(R.cpu(), ref_ops[0][0]["p5"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(p.cpu(), ref_ops[2][0]["C"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("kind,ntrans,n,d,h,ncls", [("SGC", 1, 70, 1433, 256, 7), ("SGC", 2, 909, 128, 256, 40),
                                                     ("GCN", 1, 446, 500, 256, 7)])
def test_training_gradients_through_chain(kind, ntrans, n, d, h, ncls):
    """model.train_grads (forward, nll residual, backward) recorded and run by the persistent kernel equals the
    step-by-step launches at gemm_precision 0 (both exact fp32) to summation-order rounding, at the benched shapes."""
    from graphslim_b200 import engine as _engine
    from graphslim_b200.ops import CudaOps
    K0 = CudaOps("cuda:0", precision=0)
    g = torch.Generator().manual_seed(n)
    labels = torch.sort(torch.randint(0, ncls, (n,), generator=g)).values.numpy()
    lay = _engine.ClassLayout(K0, labels, ncls)
    model = _engine.build_model(K0, kind, d, h, ncls, 2, ntrans, lay)
    W = [(torch.randn(*s, generator=g) * 0.2).cuda() for s in model.param_shapes]
    model.set_weights(W)
    X = torch.randn(n, d, generator=g).cuda()
    A = torch.rand(n, n, generator=g)
    A = ((A + A.T) / (2 * n)).cuda()
    ref = [t.clone() for t in model.train_grads(X, A)]
    out = {}
    prog = C.record(K0, [model], lambda: out.setdefault("g", model.train_grads(X, A)))
    prog.run()
    torch.cuda.synchronize()
    assert prog.n_ops >= 6 and prog.n_sync < prog.n_ops
    for a, b in zip(ref, out["g"]):
        scale = float(a.abs().max()) + 1e-30
        assert float((a.reshape(-1) - b.reshape(-1)).abs().max()) <= 2e-5 * scale


@pytest.mark.parametrize("name", ["mini_sgc1_trans", "mini_sgc2_arxiv", "mini_gcn_flickr"])
def test_inner_loop_chain_matches_stepwise(name):
    """Two epochs with the inner loop run by the persistent kernel against the step-by-step launches.  Both are exact
    fp32 at gemm_precision 0 but sum in different orders, and Adam turns last-bit differences into visible ones; the
    yardstick is what switching the step-by-step run to the 3xBF16 products does to the same trajectory."""
    from graphslim_b200 import data as gdata
    from graphslim_b200.reduction import create_reducer

    def run(chain, precision):
        args = helpers.case_args(name, device="cuda", save_init=False, progress=False, gemm_precision=precision,
                                 cuda_graphs=chain, inner_chain=chain, grouped_mn=False)
        args.epochs = 2
        raw = helpers.case_graph(name)
        helpers.seed_everything(args.seed)
        data = gdata.TransAndInd(raw, args.dataset, args.pre_norm)
        agent = create_reducer(args.method, setting=args.setting, data=data, args=args)
        agent.reduce(data, verbose=False)
        torch.cuda.synchronize()
        if chain:
            assert agent.inner.chain is not None, getattr(agent.inner, "chain_error", "chain not recorded")
            assert agent.inner.graph is None and agent.inner.replays > 0
        else:
            assert agent.inner.chain is None
        return [data.feat_syn.cpu().clone(), data.adj_syn.cpu().clone()] + [w.cpu().clone() for w in agent.inner.W]

    a, b, c = run(False, 0), run(False, 1), run(True, 0)
    for x, y, z in zip(a, b, c):
        scale = float(x.abs().max()) + 1e-30
        assert float((x - z).abs().max()) <= max(10 * float((x - y).abs().max()), 2e-3 * scale)
