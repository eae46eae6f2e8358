"""ORACLE (test infrastructure) -- the parity cases shared by the golden generator
(oracle/make_goldens.py, runs the reference itself), the oracle restatement tests
and the GPU parity tests.

Each case names the reference dataset flag (it selects the JSON hyper-parameters in
graphslim/configs/<method>/<dataset>.json, the PGE width in
models/parametrized_adj.py:11-17 and the fan-outs in dataset/loader.py:197-210), the
synthetic graph fed in its place, and the epochs to run with checkpoints disabled.
"""
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case, full Cora shape.
    "cora_sgc1": dict(
        dataset="cora", method="gcond", epochs=2, graph=dict(name="cora", seed=0),
        keep_samples=1, keep_grads=2, grad_subsample=7,
    ),
    # small transductive SGC(ntrans=1) run long enough (>10 epochs) that feat_syn moves (gcond.py:58-61)
    "mini_sgc1_trans": dict(
        dataset="cora", method="gcond", epochs=12,
        graph=dict(n=600, und_edges=1500, d=96, c=5, split=(100, 100, 200), per_class_train=20, seed=3),
        overrides=dict(outer_loop=4, inner_loop=3, lr_feat=0.01, lr_adj=0.01),
        keep_samples=2, keep_grads=3, grad_subsample=3,
    ),
    # arxiv-style: SGC ntrans=2, PGE width 256, StandardScaler features, fan-outs [10,5]
    "mini_sgc2_arxiv": dict(
        dataset="ogbn-arxiv", method="gcond", epochs=11, reduction_rate=0.05,
        graph=dict(n=1500, und_edges=9000, d=48, c=6, split=(800, 200, 500), seed=5),
        overrides=dict(outer_loop=3, inner_loop=2, hidden=64),
        keep_samples=1, keep_grads=2, grad_subsample=3,
    ),
    # flickr-style: inductive (train-induced subgraph), condense_model GCN, fan-outs [15,8]
    "mini_gcn_flickr": dict(
        dataset="flickr", method="gcond", epochs=11, reduction_rate=0.05,
        graph=dict(n=1600, und_edges=12000, d=40, c=4, split=(800, 400, 400), seed=7),
        overrides=dict(outer_loop=3, inner_loop=1, condense_model="GCN", hidden=64),
        keep_samples=1, keep_grads=2, grad_subsample=3,
    ),
    # reddit-style with reduction_rate 0.01: 5-chunk PGE with per-chunk BatchNorm statistics
    "mini_sgc1_reddit_chunked": dict(
        dataset="reddit", method="gcond", epochs=3, reduction_rate=0.01,
        graph=dict(n=6000, und_edges=60000, d=32, c=5, split=(4000, 500, 1500), seed=9),
        overrides=dict(outer_loop=2, inner_loop=1, lr_feat=0.01, lr_adj=0.01),
        keep_samples=1, keep_grads=2, grad_subsample=3,
    ),
    # GCondX: identity structure, JSON default dis_metric 'mse', condense_model GCN
    "mini_gcondx_mse": dict(
        dataset="cora", method="gcondx", epochs=3,
        graph=dict(n=600, und_edges=1500, d=96, c=5, split=(100, 100, 200), per_class_train=20, seed=3),
        overrides=dict(outer_loop=6, inner_loop=2, hidden=32),
        keep_samples=1, keep_grads=2, grad_subsample=3,
    ),
    # DosCond: one-step matching, PGE and features both stepped every outer step, no inner loop;
    # cora JSON: GCN, 'mse'
    "mini_doscond_gcn": dict(
        dataset="cora", method="doscond", epochs=4,
        graph=dict(n=600, und_edges=1500, d=96, c=5, split=(100, 100, 200), per_class_train=20, seed=3),
        overrides=dict(outer_loop=3, hidden=32),
        keep_samples=1, keep_grads=2, grad_subsample=3,
    ),
    # DosCondX: identity structure; ogbn-arxiv JSON: SGC ntrans=2, 'mse'
    "mini_doscondx_sgc2": dict(
        dataset="ogbn-arxiv", method="doscondx", epochs=3, reduction_rate=0.05,
        graph=dict(n=1200, und_edges=6000, d=32, c=4, split=(600, 200, 400), seed=11),
        overrides=dict(outer_loop=3, hidden=32),
        keep_samples=1, keep_grads=2, grad_subsample=3,
    ),
    # `--agg` init: initial features = two-hop aggregation A_hat^2 X of the selected nodes
    # (sparsification/model_free_coreset_base.py:18-27), transductive SGC ntrans=1
    "mini_sgc1_agg": dict(
        dataset="cora", method="gcond", epochs=2,
        graph=dict(n=600, und_edges=1500, d=96, c=5, split=(100, 100, 200), per_class_train=20, seed=3),
        overrides=dict(outer_loop=3, inner_loop=2, agg=True),
        keep_samples=1, keep_grads=2, grad_subsample=3,
    ),
    # 'cos' metric on SGC ntrans=2
    "mini_sgc2_cos": dict(
        dataset="ogbn-arxiv", method="gcond", epochs=2, reduction_rate=0.05,
        graph=dict(n=1200, und_edges=6000, d=32, c=4, split=(600, 200, 400), seed=11),
        overrides=dict(outer_loop=2, inner_loop=1, hidden=32, dis_metric="cos"),
        keep_samples=1, keep_grads=2, grad_subsample=3,
    ),
}
