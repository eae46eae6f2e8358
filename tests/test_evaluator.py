"""Checkpoint evaluator (graphslim_b200/evaluation.py) against fixtures produced by the UNMODIFIED reference's
GCondBase.test_with_val (oracle/make_eval_goldens.py): same condensed graph, same generator state.

* the eval model's initial parameters are bit-exact (torch CPU generator, constructor draw + initialize());
* the validation accuracy after each of the first training iterations and the best validation / test accuracies agree
  to within a few validation nodes (fp32 reassociation flips an argmax here and there).

CPU: kernels replaced by their PyTorch references (tests/emu_ops.py).  GPU: the CUDA path through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle.cases import GOLDEN_DIR
from tests import helpers

CASES = ["mini_sgc2_arxiv", "mini_gcn_flickr"]


def _run(name, K, device):
    from graphslim_b200 import data as gdata
    from graphslim_b200.evaluation import GCNEvaluator
    gold = np.load(os.path.join(GOLDEN_DIR, f"eval_{name}.npz"))
    args = helpers.case_args(name, device=device, save_init=False, progress=False)
    args.eval_epochs = int(gold["iters"])
    raw = helpers.case_graph(name)
    data = gdata.TransAndInd(raw, args.dataset, args.pre_norm)
    data.adj_syn = torch.from_numpy(gold["adj_syn"])
    data.feat_syn = torch.from_numpy(gold["feat_syn"])
    data.labels_syn = torch.from_numpy(gold["labels_syn"])
    ev = GCNEvaluator(K, data, args)
    helpers.seed_everything(args.seed + 17)
    out = []
    for run in range(gold["res"].shape[0]):
        res = ev.test_with_val()
        init = np.concatenate([w.cpu().numpy().ravel() for w in ev.last_init])
        out.append((res, init, np.array(ev.last_val_curve)))
    probe = torch.randint(0, 2**31 - 1, (4,)).numpy()
    return gold, out, probe, data


def _check(gold, out, probe, data):
    n_val = len(np.asarray(data.labels_val))
    n_test = len(np.asarray(data.labels_test))
    for run, (res, init, curve) in enumerate(out):
        np.testing.assert_array_equal(init, gold["model_init"][run])          # parameter draws: bit exact
        ref_curve = gold["val_curve"][run]
        assert curve.shape == ref_curve.shape
        # early iterations: identical state, only fp32 reassociation -> at most a couple of validation nodes apart
        assert np.abs(curve[:10] - ref_curve[:10]).max() <= 3.0 / n_val + 1e-12
        # whole trajectory: bounded drift
        assert np.abs(curve - ref_curve).max() <= 0.05
        assert abs(res[0] - gold["res"][run, 0]) <= 3.0 / n_val + 1e-12       # best validation accuracy
        assert abs(res[1] - gold["res"][run, 1]) <= 0.05 + 3.0 / n_test       # test accuracy of the selected model
    np.testing.assert_array_equal(probe, gold["torch_rng_probe"])             # total generator consumption


@pytest.mark.parametrize("name", CASES)
def test_evaluator_emulated_matches_reference_fixture(name):
    from tests.emu_ops import EmuOps
    _check(*_run(name, EmuOps("cpu"), "cpu"))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("name", CASES)
def test_evaluator_cuda_matches_reference_fixture(name, precision):
    from graphslim_b200.ops import CudaOps
    _check(*_run(name, CudaOps("cuda:0", precision=precision), "cuda"))


def test_reduce_with_checkpoints_runs_the_builtin_evaluator(tmp_path, monkeypatch):
    """GCond.reduce with a checkpoint epoch: the published condensed graph goes through intermediate_evaluation
    (run_inter_eval runs of the built-in GCN evaluator) and is saved in the reference's three-file format when the
    validation accuracy improves (gcond.py:75-78, gcond_base.py:287-324, dataset/utils.py:136-152)."""
    import logging
    from graphslim_b200 import data as gdata
    from graphslim_b200.condensation import gcond_base
    from graphslim_b200.reduction import create_reducer
    from tests.emu_ops import EmuOps
    monkeypatch.setattr(gcond_base, "_kernels", lambda device, args: EmuOps(device))
    name = "mini_sgc2_arxiv"
    args = helpers.case_args(name, save_init=False, progress=False, save_path=str(tmp_path))
    args.epochs, args.checkpoints, args.eval_epochs, args.run_inter_eval = 3, [1], 8, 2
    records = []
    handler = logging.Handler()
    handler.emit = lambda rec: records.append(rec.getMessage())
    args.logger.addHandler(handler)
    args.logger.setLevel(logging.INFO)
    raw = helpers.case_graph(name)
    helpers.seed_everything(args.seed)
    data = gdata.TransAndInd(raw, args.dataset, args.pre_norm)
    agent = create_reducer(args.method, setting=args.setting, data=data, args=args)
    try:
        agent.reduce(data, verbose=False)
    finally:
        args.logger.removeHandler(handler)
    assert 0.0 < agent.best_val <= 1.0
    assert any(m.strip().startswith("Val:") for m in records) and any(m.startswith("Test:") for m in records)
    saved = [f for _, _, fs in os.walk(tmp_path) for f in fs if f.endswith(".pt")]
    assert any(f.startswith("adj_") for f in saved) and any(f.startswith("feat_") for f in saved) and \
        any(f.startswith("label_") for f in saved)


def _directed_hub_graph(n=400, hub_out=150, seed=0):
    """A directed graph with one hub ROW (out-degree > the 64-nnz long-row threshold) whose transpose has a hub COLUMN
    instead: the row lengths of the normalised matrix and of its transpose differ."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    rows = np.concatenate([np.zeros(hub_out, dtype=np.int64), rng.integers(1, n, 600)])
    cols = np.concatenate([rng.choice(np.arange(1, n), hub_out, replace=False), rng.integers(0, n, 600)])
    return sp.coo_matrix((np.ones(rows.size, dtype=np.float32), (rows, cols)), shape=(n, n)).tocsr()


def test_normalized_csr_of_directed_graph_is_the_transpose():
    """the returned CSR is the transpose of the normalised matrix (SparseTensor(...).t()), also when only
    the VALUES are asymmetric; checked against scipy on the CPU-emulated kernels."""
    import scipy.sparse as sp
    from graphslim_b200.evaluation import normalized_csr
    from tests.emu_ops import EmuOps
    a = _directed_hub_graph()
    K = EmuOps("cpu")
    csr = normalized_csr(K, a)
    b = (sp.csr_matrix(a, dtype=np.float64) + sp.eye(a.shape[0])).tocsr()
    r = np.power(np.asarray(b.sum(1)).ravel(), -0.5)
    ref = (sp.diags(r) @ b @ sp.diags(r)).T.tocsr()
    ref.sort_indices()
    assert np.array_equal(csr.rowptr.numpy(), ref.indptr) and np.array_equal(csr.col.numpy(), ref.indices)
    np.testing.assert_allclose(csr.val.numpy(), ref.data.astype(np.float32), rtol=1e-6)
    X = torch.randn(a.shape[0], 8)
    np.testing.assert_allclose(K.spmm(csr, X).numpy(), (ref @ X.numpy().astype(np.float64)), rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_normalized_csr_directed_hub_rows_on_cuda():
    """The long-row chunk table must describe the returned (transposed) CSR: a hub column of the graph becomes a hub row
    of the transpose, which the SpMM only covers through its chunk items."""
    import scipy.sparse as sp
    from graphslim_b200.evaluation import normalized_csr
    from graphslim_b200.ops import CudaOps
    a = _directed_hub_graph().T.tocsr()           # hub column -> the returned transpose has the hub ROW
    K = CudaOps("cuda", precision=0)
    csr = normalized_csr(K, a)
    b = (sp.csr_matrix(a, dtype=np.float64) + sp.eye(a.shape[0])).tocsr()
    r = np.power(np.asarray(b.sum(1)).ravel(), -0.5)
    ref = (sp.diags(r) @ b @ sp.diags(r)).T.tocsr()
    assert int(np.diff(ref.indptr).max()) > 64 and csr.chunks is not None
    X = torch.randn(a.shape[0], 32)
    got = K.spmm(csr, X.cuda()).cpu().numpy()
    np.testing.assert_allclose(got, ref @ X.numpy().astype(np.float64), rtol=1e-4, atol=1e-5)
