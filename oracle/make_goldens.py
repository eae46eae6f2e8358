"""ORACLE (test infrastructure) -- generate tests/golden/*.npz by running the
UNMODIFIED reference GCond/GCondX from /root/reference through the import shim
(oracle/ref_shim).  Runs only in the build container; the fixtures it writes are
committed so the GPU box (no /root/reference) can check against them.

    python -m oracle.make_goldens [case ...]

What is recorded per case (all produced by reference code; hooks only observe):
  * graph CSR after the reference's normalize_adj_tensor(sparse=True): rowptr/col/val
  * labels_syn, num_class_dict order, Random-init node ids (via feat_init rows)
  * per class step: the class batch, n_id and both sampled blocks (first
    `keep_samples` outer steps in full, every step as a checksum)
  * per outer step: loss; for the first `keep_grads` steps feat_syn.grad and PGE grads
  * feat_syn / PGE parameters after the last epoch; the condense model's
    freshly initialised parameters at the start of every epoch
"""
import argparse
import hashlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from graphslim_b200 import synth  # noqa: E402
from oracle.cases import CASES, GOLDEN_DIR  # noqa: E402


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return np.frombuffer(h.digest()[:8], dtype=np.uint64)[0]


def reference_args(case, save_path):
    """Resolve `args` exactly as the reference CLI does (config.py:363-399), then apply the case overrides."""
    from graphslim import config
    argv = ["-D", case["dataset"], "-M", case["method"], "-G", "-1", "--save_path", save_path, "-E", "20",
            "-S", str(case.get("seed", 1))]
    if "reduction_rate" in case:
        argv += ["-R", str(case["reduction_rate"])]
    args = config.cli.main(args=argv, standalone_mode=False)
    for k, v in case.get("overrides", {}).items():
        setattr(args, k, v)
    args.epochs = case["epochs"]
    args.checkpoints = []
    args.verbose = False
    return args


def build_reference_data(case, args):
    from graphslim.dataset.loader import TransAndInd
    from graphslim.dataset.utils import splits
    raw = synth.make_graph(**case["graph"])
    raw = splits(raw, "fixed")
    data = TransAndInd(raw, case["dataset"], args.pre_norm)
    data.nclass = raw.num_classes
    return data


def run_case(name):
    from oracle import ref_shim
    ref_shim.install()
    from graphslim.models.base import BaseGNN
    from graphslim.reduction import create_reducer
    from graphslim.utils import seed_everything
    import graphslim.utils as gutils

    case = CASES[name]
    tmp = tempfile.mkdtemp(prefix="gs_golden_")
    args = reference_args(case, tmp)
    data = build_reference_data(case, args)
    rec = {}
    keep_samples = case.get("keep_samples", 1)
    keep_grads = case.get("keep_grads", 2)
    sub = case.get("grad_subsample", 1)

    seed_everything(args.seed)
    agent = create_reducer(args.method, setting=args.setting, data=data, args=args)
    rec["labels_syn"] = np.asarray(data.labels_syn).astype(np.int64)
    rec["class_order"] = np.array(list(agent.num_class_dict.keys()), dtype=np.int64)
    rec["class_count"] = np.array(list(agent.num_class_dict.values()), dtype=np.int64)
    rec["pge_init"] = np.concatenate([p.detach().numpy().ravel() for p in agent.pge.parameters()])[::sub]

    # ---- observers ---------------------------------------------------------------------------
    state = dict(step=0, cls_calls=0, epoch=0)
    losses, sample_digest, model_init = [], [], []

    orig_norm = gutils.normalize_adj_tensor
    import graphslim.condensation.gcond as gmod
    import graphslim.condensation.gcondx as gxmod

    def norm_spy(adj, sparse=False):
        out = orig_norm(adj, sparse=sparse)
        if sparse and "adj_rowptr" not in rec:
            rp, c, v = out.csr()
            rec["adj_rowptr"], rec["adj_col"], rec["adj_val"] = rp.numpy().copy(), c.numpy().copy(), v.numpy().copy()
        return out

    import graphslim.condensation.doscond as dmod
    import graphslim.condensation.doscondx as dxmod
    for m in (gmod, gxmod, dmod, dxmod):
        m.normalize_adj_tensor = norm_spy

    orig_sampler = data.retrieve_class_sampler

    def sampler_spy(c, adj, a, num=256):
        out = orig_sampler(c, adj, a, num)
        bs, n_id, adjs = out
        parts = [n_id.numpy()]
        for (adj_t, _, _) in adjs:
            rp, cc, vv = adj_t.csr()
            parts += [rp.numpy(), cc.numpy(), vv.numpy()]
        sample_digest.append(_digest(*parts))
        ostep = state["cls_calls"] // data.nclass
        if ostep < keep_samples:
            k = f"s{ostep}_c{c}"
            rec[k + "_bs"] = np.int64(bs)
            rec[k + "_nid"] = n_id.numpy().copy()
            for h, (adj_t, _, _) in enumerate(adjs):
                rp, cc, vv = adj_t.csr()
                rec[f"{k}_h{h}_rowptr"], rec[f"{k}_h{h}_col"], rec[f"{k}_h{h}_val"] = \
                    rp.numpy().copy(), cc.numpy().copy(), vv.numpy().copy()
        state["cls_calls"] += 1
        return out

    data.retrieve_class_sampler = sampler_spy

    orig_train_class = agent.train_class

    def train_class_spy(*a, **k):
        loss = orig_train_class(*a, **k)
        losses.append(float(loss.item()))
        return loss

    agent.train_class = train_class_spy

    def grads_spy(opt_step):
        def wrapped(*a, **k):
            s = state["step"]
            if s < keep_grads:
                rec[f"g{s}_feat"] = agent.feat_syn.grad.detach().numpy().copy()[:, ::sub]
                pg = [p.grad.detach().numpy().ravel() if p.grad is not None else np.zeros(p.numel(), np.float32)
                      for p in agent.pge.parameters()]
                rec[f"g{s}_pge"] = np.concatenate(pg)[::sub]
            if s == 0:
                rec["feat_init"] = agent.feat_syn.detach().numpy().copy()[:, ::sub]
                if agent.adj_syn is not None:
                    rec["adj_syn_norm0"] = agent.adj_syn.detach().numpy().copy()
            state["step"] += 1
            return opt_step(*a, **k)
        return wrapped

    agent.optimizer_feat.step = grads_spy(agent.optimizer_feat.step)
    if args.method not in ("doscond", "doscondx"):      # DosCond steps both optimisers every outer step: count once
        agent.optimizer_pge.step = grads_spy(agent.optimizer_pge.step)

    orig_init = BaseGNN.initialize

    def init_spy(self):
        orig_init(self)
        model_init.append(np.concatenate([p.detach().numpy().ravel() for p in self.parameters()])[::sub])

    BaseGNN.initialize = init_spy
    try:
        agent.reduce(data, verbose=False)
    finally:
        BaseGNN.initialize = orig_init

    rec["losses"] = np.array(losses, dtype=np.float64)
    rec["sample_digest"] = np.array(sample_digest, dtype=np.uint64)
    rec["model_init"] = np.stack(model_init)
    rec["feat_final"] = agent.feat_syn.detach().numpy().copy()[:, ::sub]
    rec["pge_final"] = np.concatenate([p.detach().numpy().ravel() for p in agent.pge.parameters()])[::sub]
    if args.method == "gcond":
        with torch.no_grad():
            pass
    # RNG positions at exit pin the total stream consumption
    rec["np_rng_probe"] = np.random.randint(0, 2**31 - 1, size=4).astype(np.int64)
    rec["torch_rng_probe"] = torch.randint(0, 2**31 - 1, (4,)).numpy()
    rec["grad_subsample"] = np.int64(sub)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    out = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(out, **rec)
    print(f"[golden] {name}: {len(losses)} outer steps, first losses {losses[:3]}, wrote {out} "
          f"({os.path.getsize(out) / 1e6:.2f} MB)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=list(CASES))
    ns = ap.parse_args()
    for name in ns.cases:
        run_case(name)


if __name__ == "__main__":
    main()
