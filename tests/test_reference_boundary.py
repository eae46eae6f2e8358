"""The drop-in boundary exercised with the REFERENCE'S OWN objects (CPU, kernels replaced by their PyTorch references):

  * `graphslim_b200.condensation.gcond.GCond` constructed with the reference's real `TransAndInd` data object and the
    `args` namespace its click CLI produces (graphslim/config.py:363-399), `reduce(data, verbose=True)` printing the
    `verbose_time_memory` lines, losses equal to the fixture the unmodified reference produced on the same inputs;
  * the reference's `graphslim/train_all.py:19-38` `main()` run end to end with its registry entry for `gcond` pointed at
    this package: flags parsed by the reference, reducer resolved by name, checkpoints
    evaluated and saved through `save_reduced`, files readable by the reference's own `get_syn_data`.

Needs the reference sources (/root/reference, or the copy oracle/stage_ref.py stages under oracle/_ref); skipped otherwise.
"""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_root():
    for cand in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        if os.path.isfile(os.path.join(cand, "graphslim", "condensation", "gcond.py")):
            return cand
    return None


pytestmark = pytest.mark.skipif(_reference_root() is None, reason="reference sources not available")


@pytest.fixture(scope="module")
def ref():
    os.environ.setdefault("GRAPHSLIM_REFERENCE_ROOT", _reference_root())
    from oracle import ref_shim
    ref_shim.install()
    from oracle import make_goldens as MG
    return MG


@pytest.fixture
def emulated(monkeypatch):
    from graphslim_b200.condensation import gcond_base
    from tests.emu_ops import EmuOps
    monkeypatch.setattr(gcond_base, "_kernels", lambda device, args: EmuOps("cpu"))


def test_product_reducer_consumes_reference_data_and_cli_args(ref, emulated, capsys):
    from graphslim.utils import seed_everything as ref_seed
    from graphslim_b200.condensation.gcond import GCond
    from oracle.cases import CASES
    from tests import helpers
    name = "mini_sgc2_arxiv"
    case = dict(CASES[name], epochs=2)
    tmp = tempfile.mkdtemp(prefix="gs_boundary_")
    args = ref.reference_args(case, tmp)                 # the reference's own CLI -> Obj namespace
    data = ref.build_reference_data(case, args)          # the reference's own TransAndInd
    assert type(data).__module__ == "graphslim.dataset.loader" and type(args).__module__.startswith("graphslim")
    ref_seed(args.seed)
    agent = GCond(setting=args.setting, data=data, args=args)
    losses = []
    agent.trace = lambda kind, **kw: losses.append(float(kw["loss"].item())) if kind == "grads" else None
    out = agent.reduce(data, verbose=True)
    assert out is data                                   # same object back, mutated in place (gcond.py:81)
    printed = capsys.readouterr().out
    assert "Function Time:" in printed and "Original graph:" in printed and "Condensed graph:" in printed
    gold = helpers.golden(name)
    n = len(losses)
    assert n == 2 * args.outer_loop
    np.testing.assert_allclose(losses[:2], gold["losses"][:2], rtol=1e-4)
    np.testing.assert_allclose(losses, gold["losses"][:n], rtol=helpers.PARITY_TOL[name][0][0])
    assert np.array_equal(np.asarray(data.labels_syn), gold["labels_syn"])
    assert tuple(data.feat_syn.shape) == (gold["labels_syn"].shape[0], data.feat_train.shape[1])
    assert tuple(data.adj_syn.shape) == (gold["labels_syn"].shape[0],) * 2
    # the Random init was saved where the reference saves it (reduced_graph/random/*.pt, coalesced sparse adjacency)
    init_dir = os.path.join(tmp, "reduced_graph", "random")
    tag = f"{args.dataset}_{args.reduction_rate}_{args.seed}.pt"
    adj0 = torch.load(os.path.join(init_dir, "adj_" + tag))
    assert adj0.layout == torch.sparse_coo and os.path.exists(os.path.join(init_dir, "feat_" + tag))


def test_reference_train_all_main_with_patched_registry(ref, emulated, monkeypatch, capsys):
    """graphslim/train_all.py main(): get_args -> get_dataset -> seed -> create_reducer('gcond') -> reduce -> evaluate."""
    import graphslim.train_all as train_all
    from graphslim.dataset.utils import get_syn_data
    from graphslim.reduction import registry
    from oracle.cases import CASES
    case = CASES["mini_sgc1_trans"]
    tmp = tempfile.mkdtemp(prefix="gs_trainall_")
    # the registry entry of `gcond` points at this package
    monkeypatch.setitem(registry._METHODS, "gcond",
                        registry.MethodSpec("gcond", "condensation", "graphslim_b200.condensation.gcond", "GCond"))
    seen = {}

    def fake_get_dataset(name, args, load_path=None):
        data = ref.build_reference_data(case, args)
        seen["data"] = data
        return data

    class _Evaluator:
        def __init__(self, args):
            seen["args"] = args

        def evaluate(self, reduced, model_type="GCN"):
            seen["reduced"] = reduced
            return 0.0, 0.0

    monkeypatch.setattr(train_all, "get_dataset", fake_get_dataset)
    monkeypatch.setattr(train_all, "Evaluator", _Evaluator)
    monkeypatch.setattr(sys, "argv", ["train_all.py", "-D", "cora", "-M", "gcond", "-G", "-1", "-E", "10", "-S", "1",
                                      "--save_path", tmp, "--outer_loop", "2", "--inner_loop", "1", "--eval_epochs", "6",
                                      "-V"])
    train_all.main()
    args, data = seen["args"], seen["data"]
    assert seen["reduced"] is data and args.method == "gcond" and args.checkpoints       # reference flag plumbing intact
    printed = capsys.readouterr().out
    assert "Function Time:" in printed                    # reduce(..., verbose=args.verbose) through the decorator
    n_syn = int(np.asarray(data.labels_syn).shape[0])
    assert tuple(data.feat_syn.shape) == (n_syn, data.feat_train.shape[1]) and tuple(data.adj_syn.shape) == (n_syn, n_syn)
    # the checkpoint evaluation saved the best condensed graph; the reference's own loader reads it back
    feat, adj, lab = get_syn_data(data, args, model_type="GCN")
    assert tuple(feat.shape) == (n_syn, data.feat_train.shape[1]) and tuple(adj.shape) == (n_syn, n_syn)
    assert int(lab.shape[0]) == n_syn
