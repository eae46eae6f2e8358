"""Argument namespace for the GCond / GCondX path.

Mirrors how the reference resolves ``args`` (graphslim/config.py:363-399): click defaults
(:260-361), then the per-method JSON hyper-parameters (graphslim/configs/<method>/<dataset>.json
via ``method_config`` :240-257), then ``setting_config`` (:209-236), then explicit overrides
(the reference re-applies command-line values last, :384-386).  Only the flags the path reads
are kept.
"""
import logging
from types import SimpleNamespace

# click defaults of the flags read on the path (config.py:260-361)
_DEFAULTS = dict(
    dataset="cora", method="gcond", gpu_id=0, setting=None, split="fixed", hidden=256, condense_model="SGC",
    epochs=1000, agg=False, multi_label=False, dis_metric="ours", lr_adj=1e-4, lr_feat=1e-4, optim="Adam",
    threshold=0.0, dropout=0.0, ntrans=1, with_bn=False, save_path="../checkpoints", load_path="./data",
    with_structure=1, lr=0.01, weight_decay=0.0, pre_norm=True, outer_loop=10, inner_loop=1,
    reduction_rate=-1.0, seed=1, nlayers=2, verbose=False, soft_label=0, init="random", eval_epochs=300,
    eval_model="GCN", run_inter_eval=5, eval_interval=100, alpha=0.1, attack=None, run_reduction=3,
    # not a reference flag: arithmetic of the large dense products (0 fp32 SIMT, 1 tcgen05 3xBF16 split, 2 tcgen05 BF16)
    gemm_precision=1,
)

_CITATION = dict(lr_feat=1e-4, lr_adj=1e-4, pre_norm=True, dis_metric="ours", outer_loop=20, inner_loop=15,
                 threshold=0.05, condense_model="SGC", ntrans=1)
_X_SMALL = dict(lr_feat=0.01, lr_adj=0.01, dis_metric="mse", pre_norm=True, outer_loop=10, condense_model="GCN")

# graphslim/configs/gcond/*.json and graphslim/configs/gcondx/*.json
METHOD_CONFIGS = {
    "gcond": {
        "cora": _CITATION, "citeseer": _CITATION, "pubmed": _CITATION, "amazon": _CITATION, "yelp": _CITATION,
        "flickr": dict(lr_feat=0.005, lr_adj=0.005, outer_loop=10, inner_loop=1, dis_metric="ours", threshold=0.01,
                       condense_model="SGC", ntrans=2),
        "ogbn-arxiv": dict(lr_feat=0.01, lr_adj=0.01, outer_loop=20, inner_loop=3, dis_metric="ours",
                           threshold=0.01, condense_model="SGC", ntrans=2, epochs=600),
        "reddit": dict(lr_feat=0.1, lr_adj=0.1, outer_loop=10, inner_loop=1, dis_metric="ours", threshold=0.01,
                       condense_model="SGC", ntrans=1, epochs=1000),
    },
    "gcondx": {
        "cora": _X_SMALL, "citeseer": _X_SMALL, "pubmed": _X_SMALL, "amazon": _X_SMALL, "yelp": _X_SMALL,
        "flickr": dict(lr_feat=0.01, lr_adj=0.01, dis_metric="mse", outer_loop=10, condense_model="GCN"),
        "ogbn-arxiv": dict(lr_feat=0.1, lr_adj=0.1, dis_metric="mse", outer_loop=5, condense_model="SGC", ntrans=2),
        "reddit": dict(lr_feat=0.1, lr_adj=0.1, outer_loop=10, inner_loop=1, dis_metric="ours", threshold=0.01,
                       condense_model="SGC", ntrans=1, epochs=400),
    },
}

# graphslim/configs/doscond/*.json and graphslim/configs/doscondx/*.json
_DOS_SMALL = dict(lr_feat=0.01, lr_adj=0.01, dis_metric="mse", outer_loop=3, threshold=0.05, condense_model="GCN")
_DOSX_SMALL = dict(lr_feat=0.01, lr_adj=0.01, dis_metric="mse", pre_norm=True, outer_loop=10, condense_model="GCN")
METHOD_CONFIGS["doscond"] = {
    "cora": _DOS_SMALL, "citeseer": dict(_DOS_SMALL, outer_loop=5), "pubmed": dict(_DOS_SMALL, outer_loop=5),
    "flickr": dict(lr_feat=5e-3, lr_adj=5e-3, dis_metric="mse", outer_loop=10, threshold=0.01, condense_model="GCN"),
    "ogbn-arxiv": dict(lr_feat=0.01, lr_adj=0.02, dis_metric="ours", outer_loop=5, threshold=0.01,
                       condense_model="SGC", ntrans=1),
    "reddit": dict(lr_feat=0.1, lr_adj=0.1, dis_metric="ours", outer_loop=10, threshold=0.01, condense_model="GCN",
                   epochs=1000),
}
METHOD_CONFIGS["doscondx"] = {
    "cora": _DOSX_SMALL, "citeseer": _DOSX_SMALL, "pubmed": _DOSX_SMALL,
    "flickr": dict(lr_feat=0.01, lr_adj=0.01, dis_metric="mse", outer_loop=10, condense_model="GCN"),
    "ogbn-arxiv": dict(lr_feat=0.1, lr_adj=0.1, dis_metric="mse", outer_loop=5, condense_model="SGC", ntrans=2),
    "reddit": dict(lr_feat=0.1, lr_adj=0.1, dis_metric="mse", outer_loop=5, condense_model="GCN"),
}

# config.py:210-220 (the later 'pubmed' key wins)
_REPRESENTATIVE_RATE = {"cora": 0.5, "citeseer": 0.5, "pubmed": 0.1, "flickr": 0.01, "reddit": 0.001,
                        "ogbn-arxiv": 0.01, "yelp": 0.001, "amazon": 0.002}


def setting_config(args):
    """config.py:209-236."""
    if args.reduction_rate == -1:
        args.reduction_rate = _REPRESENTATIVE_RATE[args.dataset]
    if args.dataset in ("cora", "citeseer", "pubmed", "ogbn-arxiv"):
        args.setting = "trans"
    if args.dataset in ("flickr", "reddit", "amazon", "yelp"):
        args.setting = "ind"
    args.metric = "f1_macro" if args.dataset in ("yelp", "amazon") else "accuracy"
    args.run_inter_eval = 3
    args.eval_interval = args.epochs // 10
    if args.eval_interval > 0:
        args.checkpoints = list(range(-1, args.epochs + 1, args.eval_interval))
    else:
        # the reference raises ValueError here (range step 0, efficiency.md:31); we keep running
        # with no checkpoints instead, which is what every timing run wants anyway.
        args.checkpoints = []
    args.eval_epochs = 300
    return args


def make_args(dataset="cora", method="gcond", gpu_id=0, **overrides):
    """Resolve the namespace the reducers read.  ``overrides`` play the role of explicit CLI flags."""
    d = dict(_DEFAULTS)
    d.update(dataset=dataset, method=method, gpu_id=gpu_id)
    args = SimpleNamespace(**d)
    args.device = f"cuda:{gpu_id}" if gpu_id >= 0 else "cpu"
    conf = METHOD_CONFIGS.get(method, {}).get(dataset)
    if conf is None:
        print("No config file found or error in json format, please use method_config(args)")
    else:
        for k, v in conf.items():
            setattr(args, k, v)
    if "epochs" in overrides:          # setting_config derives checkpoints from the final epochs
        args.epochs = overrides["epochs"]
    args = setting_config(args)
    for k, v in overrides.items():
        setattr(args, k, v)
    args.logger = logging.getLogger("graphslim_b200")
    return args
