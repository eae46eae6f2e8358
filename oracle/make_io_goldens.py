"""ORACLE (test infrastructure) -- fixtures for the on-disk result interface: the reference's
own `sparsify` / `get_syn_data` (graphslim/dataset/utils.py:8-66,258-296) run through the import shim on a seeded
condensed graph written with the reference's `save_reduced`, for several (method, evaluator) pairs.  Also copies one
of the reference's own saved results (interface/reduced_graph/gcond/adj_cora_0.5_1.pt, a data file) next to the
fixtures so the loader is checked against a file the reference wrote.  Build container only.

    python -m oracle.make_io_goldens
"""
import logging
import os
import shutil
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle.cases import GOLDEN_DIR  # noqa: E402

PAIRS = [("gcond", "GCN", 0.05), ("gcond", "GAT", 0.05), ("gcond", "MLP", 0.05), ("doscond", "SGC", 0.01),
         ("gcondx", "GCN", 0.05), ("gcondx", "GAT", 0.05)]


def main():
    from oracle import ref_shim
    ref_shim.install()
    from graphslim.dataset.utils import get_syn_data, save_reduced
    gen = torch.Generator().manual_seed(123)
    n, d = 30, 8
    adj = torch.rand(n, n, generator=gen)
    adj = ((adj + adj.T) / 2) * (1 - torch.eye(n))
    feat = torch.randn(n, d, generator=gen)
    labels = torch.randint(0, 4, (n,), generator=gen)
    rec = dict(adj=adj.numpy(), feat=feat.numpy(), labels=labels.numpy())
    data = SimpleNamespace(labels_train=torch.zeros(100, dtype=torch.long), feat_train=torch.zeros(100, d),
                           feat_full=torch.zeros(200, d))
    for method, model_type, thr in PAIRS:
        tmp = tempfile.mkdtemp(prefix="gs_io_golden_")
        args = SimpleNamespace(save_path=tmp, method=method, dataset="cora", reduction_rate=0.5, seed=1, attack=None,
                               device="cpu", setting="trans", threshold=thr, logger=logging.getLogger("io_golden"))
        save_reduced(adj.clone(), feat.clone(), labels.clone(), args)
        f, a, l = get_syn_data(data, args, model_type)
        rec[f"{method}_{model_type}_adj"] = a.numpy().copy()
        assert torch.equal(f, feat) and torch.equal(l, labels)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "io_sparsify.npz"), **rec)
    dst = os.path.join(GOLDEN_DIR, "ref_saved", "reduced_graph", "gcond")
    os.makedirs(dst, exist_ok=True)
    shutil.copy("/root/reference/interface/reduced_graph/gcond/adj_cora_0.5_1.pt", dst)
    print("[io golden] wrote io_sparsify.npz and ref_saved/reduced_graph/gcond/adj_cora_0.5_1.pt")


if __name__ == "__main__":
    main()
