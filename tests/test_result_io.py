"""On-disk result interface (graphslim_b200/dataset_utils.py) against the reference's own `save_reduced` /
`get_syn_data` / `sparsify` (fixtures from oracle/make_io_goldens.py) and against a file the reference wrote."""
import logging
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from graphslim_b200 import dataset_utils as du
from oracle.cases import GOLDEN_DIR

PAIRS = [("gcond", "GCN", 0.05), ("gcond", "GAT", 0.05), ("gcond", "MLP", 0.05), ("doscond", "SGC", 0.01),
         ("gcondx", "GCN", 0.05), ("gcondx", "GAT", 0.05)]


def _args(tmp, method, thr, **kw):
    base = dict(save_path=str(tmp), method=method, dataset="cora", reduction_rate=0.5, seed=1, attack=None, device="cpu",
                setting="trans", threshold=thr, logger=logging.getLogger("test_io"))
    base.update(kw)
    return SimpleNamespace(**base)


@pytest.mark.parametrize("method,model_type,thr", PAIRS)
def test_save_load_sparsify_match_reference(tmp_path, method, model_type, thr):
    gold = np.load(os.path.join(GOLDEN_DIR, "io_sparsify.npz"))
    adj, feat, labels = (torch.from_numpy(gold[k]) for k in ("adj", "feat", "labels"))
    args = _args(tmp_path, method, thr)
    du.save_reduced(adj.clone(), feat.clone(), labels.clone(), args)
    for kind in ("adj", "feat", "label"):             # the reference's three-file layout
        assert os.path.exists(tmp_path / "reduced_graph" / method / f"{kind}_cora_0.5_1.pt")
    data = SimpleNamespace(labels_train=torch.zeros(100, dtype=torch.long), feat_train=torch.zeros(100, 8),
                           feat_full=torch.zeros(200, 8))
    f, a, l = du.get_syn_data(data, args, model_type)
    assert torch.equal(f, feat) and torch.equal(l, labels)
    np.testing.assert_array_equal(a.numpy(), gold[f"{method}_{model_type}_adj"])


def test_load_reads_a_file_written_by_the_reference_and_falls_back_for_the_rest(capsys):
    """interface/reduced_graph/gcond/adj_cora_0.5_1.pt of the reference (70 x 70 fp32); no feat / label file next to
    it, so those fall back to the original graph's, as dataset/utils.py:220-243 does."""
    args = _args(os.path.join(GOLDEN_DIR, "ref_saved"), "gcond", 0.05)
    data = SimpleNamespace(labels_train=torch.arange(140), feat_train=torch.zeros(140, 5), feat_full=torch.ones(2708, 5))
    adj, feat, labels = du.load_reduced(args, data)
    assert adj.shape == (70, 70) and adj.dtype == torch.float32
    assert float(adj.diagonal().abs().max()) == 0.0 and 0.0 <= float(adj.min()) and float(adj.max()) <= 1.0
    assert feat is data.feat_full and labels is data.labels_train
    out = capsys.readouterr().out
    assert "find no feat" in out and "find no label" in out


def test_nothing_saved_means_original_graph(tmp_path):
    args = _args(tmp_path, "gcond", 0.05, setting="ind")
    data = SimpleNamespace(labels_train=torch.arange(12), feat_train=torch.zeros(12, 3), feat_full=torch.ones(30, 3))
    f, a, l = du.get_syn_data(data, args, "GCN")
    assert f is data.feat_train and l is data.labels_train and torch.equal(a, torch.eye(12))
