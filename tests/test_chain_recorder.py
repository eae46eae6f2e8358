"""graphslim_b200/chain.py on CPU: recording the condense-model training step of every engine (SGC1 / SGC2 / GCN2)
yields a program whose plain-PyTorch interpretation equals the step-by-step run, barrier flags are set wherever an
operation depends on the ones before it, and calls the recorder does not know refuse to record."""
import pytest
import torch

from graphslim_b200 import chain as C
from graphslim_b200 import engine as _engine
from graphslim_b200.condensation.gcond_base import InnerLoop
from tests.chain_interp import interpret
from tests.emu_ops import EmuOps


def make(kind, ntrans, n=37, d=24, h=16, ncls=5, seed=0):
    K = EmuOps("cpu")
    g = torch.Generator().manual_seed(seed)
    labels = torch.sort(torch.randint(0, ncls, (n,), generator=g)).values.numpy()
    labels[:ncls] = range(ncls)
    labels.sort()
    lay = _engine.ClassLayout(K, labels, ncls)
    model = _engine.build_model(K, kind, d, h, ncls, 2, ntrans, lay)
    feat = torch.randn(n, d, generator=g)
    A = torch.rand(n, n, generator=g)
    A = (A + A.T) / (2 * n)
    loop = InnerLoop(K, model, feat, n, 0.01, 8, use_graph=False)
    W = [torch.randn(*s, generator=g) * 0.3 for s in model.param_shapes]
    loop.begin_epoch(W)
    loop.set_adj(A)
    return K, model, loop


@pytest.mark.parametrize("kind,ntrans", [("SGC", 1), ("SGC", 2), ("GCN", 1)])
def test_recorded_step_equals_stepwise(kind, ntrans):
    K, model, loop = make(kind, ntrans)
    K2, model2, loop2 = make(kind, ntrans)
    for _ in range(3):
        loop._one_step()                                   # reference: three eager steps
    rec = C.ChainRecorder(K2)
    saved = (loop2.K, model2.K)
    loop2.K = model2.K = rec
    try:
        loop2._one_step()                                  # records, computes nothing
    finally:
        loop2.K, model2.K = saved
    assert all(torch.equal(a, b) for a, b in zip(loop2.W, make(kind, ntrans)[2].W))     # weights untouched so far
    kinds = [f["kind"] for f, _, _ in rec.ops]
    assert kinds.count(C.ADAM_TABLE) == len(loop2.W) and kinds[-1] == C.COUNTER_ADD and C.GEMM in kinds
    for _ in range(3):
        interpret(rec.ops)
    for a, b in zip(loop.W, loop2.W):
        torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)
    assert int(loop2.step_dev.item()) == 3


def test_barrier_flags_follow_dependencies():
    K = EmuOps("cpu")
    rec = C.ChainRecorder(K)
    g = torch.Generator().manual_seed(1)
    X, W1, W2 = torch.randn(40, 8, generator=g), torch.randn(8, 6, generator=g), torch.randn(8, 5, generator=g)
    H = rec.gemm(X, W1)                     # 0
    U = rec.gemm(X, W2)                     # 1: independent of 0
    s = rec.colsum(H)                       # 2: reads 0's output
    t = rec.colsum(U)                       # 3: no new dependency since the barrier before 2
    rec.gemm(X, W1, out=H)                  # 4: overwrites what 2 read
    rec.gemm(H, W1, tb=True, out=X[:, :8])  # 5: reads 4's output, overwrites an input of 0/1/4
    live_r, live_w, flags = [], [], []
    for i, (f, reads, writes) in enumerate(rec.ops):
        dep = i > 0 and (any(C._overlap(r, w) for r in reads for w in live_w)
                         or any(C._overlap(a, b) for a in writes for b in live_w + live_r))
        flags.append(int(dep))
        if dep:
            live_r, live_w = [], []
        live_r += reads
        live_w += writes
    assert flags == [0, 0, 1, 0, 1, 1]
    import struct
    prog = rec.program()                                   # packs the same flags into the gs_chain_op array
    raw = prog.ops_dev.cpu().numpy().tobytes()
    got = [struct.unpack_from("<8i", raw, i * C._OP.size)[7] for i in range(prog.n_ops)]
    assert got == flags and prog.n_sync == sum(flags) and C._OP.size == 144


def test_unknown_calls_refuse_to_record():
    rec = C.ChainRecorder(EmuOps("cpu"))
    with pytest.raises(C.ChainUnsupported):
        rec.spmm(None, None)
    with pytest.raises(C.ChainUnsupported):
        rec.softmax_residual(torch.zeros(2, 2), torch.zeros(2, dtype=torch.int32), None, want_nll=True)
    with pytest.raises(C.ChainUnsupported):
        rec.program()
