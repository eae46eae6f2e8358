"""ORACLE (test infrastructure) -- fixtures for the checkpoint evaluator.

Runs the UNMODIFIED reference from /root/reference through the import shim: GCond.reduce for the case's epochs, then
GCondBase.test_with_val (graphslim/condensation/gcond_base.py:326-358 -> BaseGNN.fit_with_val / test,
graphslim/models/base.py:80-225) `runs` times on the condensed graph, exactly as intermediate_evaluation does
(:287-324).  Recorded: the condensed graph the evaluator saw, the RNG state it started from, and per run the eval
model's initial parameters, the validation accuracy of every training iteration, the best validation accuracy and
the test accuracy.  Build container only; the .npz files are committed.

    python -m oracle.make_eval_goldens [case ...]
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle.cases import CASES, GOLDEN_DIR  # noqa: E402
from oracle.make_goldens import build_reference_data, reference_args  # noqa: E402

EVAL_CASES = {"mini_sgc2_arxiv": dict(runs=2, iters=60), "mini_gcn_flickr": dict(runs=2, iters=60)}


def run_case(name):
    from oracle import ref_shim
    ref_shim.install()
    from graphslim.models.base import BaseGNN
    from graphslim.reduction import create_reducer
    from graphslim.utils import seed_everything
    import graphslim.utils as gutils

    case, spec = CASES[name], EVAL_CASES[name]
    args = reference_args(case, tempfile.mkdtemp(prefix="gs_eval_golden_"))
    args.eval_epochs = spec["iters"]
    data = build_reference_data(case, args)
    seed_everything(args.seed)
    agent = create_reducer(args.method, setting=args.setting, data=data, args=args)
    agent.reduce(data, verbose=False)
    # what the checkpoint branch of gcond.py:75-78 publishes before it calls intermediate_evaluation
    with torch.no_grad():
        adj = agent.pge.inference(agent.feat_syn.detach())
    data.adj_syn, data.feat_syn = adj.detach(), agent.feat_syn.detach()
    data.labels_syn = torch.as_tensor(np.asarray(agent.labels_syn if hasattr(agent, "labels_syn") else data.labels_syn)).long()
    rec = dict(adj_syn=data.adj_syn.detach().numpy().copy(), feat_syn=data.feat_syn.detach().numpy().copy(),
               labels_syn=np.asarray(data.labels_syn).astype(np.int64), iters=np.int64(spec["iters"]),
               hidden=np.int64(args.hidden), nlayers=np.int64(args.nlayers), lr=np.float64(args.lr))
    seed_everything(args.seed + 17)                        # the evaluator starts from a known generator state
    inits, val_curves = [], []
    orig_init, orig_metric = BaseGNN.initialize, args.metric

    def init_spy(self):
        orig_init(self)
        inits.append(np.concatenate([p.detach().numpy().ravel() for p in self.parameters()]))
        val_curves.append([])

    def metric_spy(output, labels):
        acc = orig_metric(output, labels)
        val_curves[-1].append(float(acc))
        return acc

    BaseGNN.initialize = init_spy
    args.metric = metric_spy
    res = []
    try:
        for _ in range(spec["runs"]):
            res.append(agent.test_with_val(verbose=False, setting=args.setting, iters=args.eval_epochs))
    finally:
        BaseGNN.initialize = orig_init
        args.metric = orig_metric
    rec["res"] = np.array(res, dtype=np.float64)                       # (runs, 2): best val acc, test acc
    rec["model_init"] = np.stack(inits)
    # the last metric call of a run is the test accuracy (BaseGNN.test), the ones before are per-iteration validation
    rec["val_curve"] = np.array([c[:spec["iters"]] for c in val_curves], dtype=np.float64)
    rec["torch_rng_probe"] = torch.randint(0, 2**31 - 1, (4,)).numpy()
    out = os.path.join(GOLDEN_DIR, f"eval_{name}.npz")
    np.savez_compressed(out, **rec)
    print(f"[eval golden] {name}: res {rec['res'].tolist()}, wrote {out} ({os.path.getsize(out) / 1e3:.1f} kB)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=list(EVAL_CASES))
    for name in ap.parse_args().cases:
        run_case(name)


if __name__ == "__main__":
    main()
