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


# ---------------------------------------------------------------------------------------------------------------
# Parity bounds of the fixture tests.  first_tol: loss / gradients of the first outer steps (identical state):
#   gemm_precision 0 -> 1e-4, the north-star fp32 bound;
#   gemm_precision 1 -> 2e-3, THE stated bound of the tcgen05 3xBF16 stage (also DESIGN.md section 5 and
#   __graft_entry__.smoke); measured on B200: <= 8.3e-4 (Cora shape, K = 1433), <= 6.3e-4 elsewhere.
# Later steps ride an Adam trajectory (g/sqrt(v) turns rounding noise on near-zero gradient entries into +-lr moves), so
# their bounds are per case: (trajectory loss, later-step gradients, final features as relative Frobenius error), each
# <= 3.5x the largest error measured for that case on B200 and on the CPU-emulated run (profiles/r2_parity_fixtures.json),
# with floors of 1e-5 / 1e-4 (2e-3 at precision 1) / 1e-5 where the measurement is at rounding level.
FIRST_TOL = {0: 1e-4, 1: 2e-3, 2: 3e-2}
PARITY_TOL = {
    "cora_sgc1": {0: (1e-5, 3.5e-3, None), 1: (1e-5, 7e-3, None)},
    "mini_sgc1_trans": {0: (5e-2, 1e-4, 0.3), 1: (0.12, 2.5e-2, 0.5)},
    "mini_sgc2_arxiv": {0: (3.5e-3, 1e-4, 4e-3), 1: (5e-2, 2.5e-2, 2e-2)},
    "mini_gcn_flickr": {0: (1e-2, 3e-3, 5e-3), 1: (1.5e-2, 7e-3, 6e-3)},
    "mini_sgc1_reddit_chunked": {0: (1e-5, 1e-4, 1e-5), 1: (6e-4, 2e-2, 1e-5)},
    "mini_gcondx_mse": {0: (5e-2, 1e-4, 3e-2), 1: (5e-2, 2e-3, 3e-2)},
    # step 6 of this case sits on a kink: a 1e-6 relative perturbation of the dense products moves its loss by 1.7e-2
    # (benchmarks/fixture_sensitivity_probe.py, CPU, any kernel), every other step by <= 5e-4
    "mini_doscond_gcn": {0: (5e-2, 6e-3, 2e-2), 1: (5e-2, 6e-3, 2.5e-2)},
    "mini_doscondx_sgc2": {0: (1e-5, 1e-4, 1e-5), 1: (1e-5, 2e-3, 1e-5)},
    "mini_sgc2_cos": {0: (1e-5, 1e-4, 1e-5), 1: (1e-5, 2e-3, 1e-5)},
    "mini_sgc1_agg": {0: (1e-4, 1e-4, 1e-4), 1: (2e-3, 2e-3, 1e-3)},
}


def parity_tol(name, precision):
    traj, later, feat = PARITY_TOL[name][precision]
    out = dict(first_tol=FIRST_TOL[precision], traj_tol=max(traj, FIRST_TOL[precision] if precision else traj),
               later_tol=later)
    if feat is not None:
        out["feat_tol"] = feat
    return out
