"""Shared helpers for the parity tests: case -> (raw graph, args, golden)."""
import os

import numpy as np

from graphslim_b200 import config, synth
from oracle.cases import CASES, GOLDEN_DIR


def case_args(name, device="cpu", **extra):
    case = CASES[name]
    over = dict(case.get("overrides", {}))
    if "reduction_rate" in case:
        over["reduction_rate"] = case["reduction_rate"]
    over.update(epochs=case["epochs"], seed=case.get("seed", 1), save_path=extra.pop("save_path", "/tmp/gs_b200_ckpt"))
    over.update(extra)
    args = config.make_args(dataset=case["dataset"], method=case["method"], gpu_id=-1 if device == "cpu" else 0,
                            **over)
    args.checkpoints = []
    args.verbose = False
    return args


def case_graph(name):
    return synth.make_graph(**CASES[name]["graph"])


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))


def seed_everything(seed):
    """graphslim/utils.py:86-91."""
    import random
    import torch
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
