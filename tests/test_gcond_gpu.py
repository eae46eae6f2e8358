"""End-to-end parity of the CUDA path: GCond / GCondX on cuda:0 (through the C ABI) against the fixtures produced
by the unmodified reference, and against the oracle restatement run in-process on the same seeded inputs.

Index work (labels_syn, A_hat CSR + values, class batches, sampled blocks, parameter draws, RNG consumption) is
bit exact; losses / gradients of the first outer steps within 1e-4 relative (fp32 everywhere, north-star bound);
multi-epoch trajectories carry per-case bounds (<= 3.5x the measured errors) because Adam's g/sqrt(v) amplifies fp32
reassociation noise (observed equally between two CPU runs of the reference that differ only in matmul blocking)."""
import numpy as np
import pytest
import torch

from oracle.cases import CASES
from tests import helpers
from tests.test_engine_emulated import check_against_golden, run_case

pytestmark = pytest.mark.gpu


# Bounds: tests/helpers.py (FIRST_TOL: 1e-4 at gemm_precision 0, the stated 2e-3 of the tcgen05 3xBF16 stage at
# precision 1; PARITY_TOL: per-case trajectory bounds <= 3.5x the measured errors).
@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("name", [c for c in CASES if c != "cora_sgc1"])
def test_gcond_cuda_matches_reference_fixture(name, precision):
    check_against_golden(*run_case(name, device="cuda", gemm_precision=precision), **helpers.parity_tol(name, precision))


@pytest.mark.parametrize("precision", [0, 1])
def test_gcond_cuda_cora_shape_first_epoch(precision):
    """BASELINE configs[0] at full Cora shape (2,708 nodes / 1,433 feats / N'=70), first epoch."""
    check_against_golden(*run_case("cora_sgc1", epochs=1, device="cuda", gemm_precision=precision),
                         **helpers.parity_tol("cora_sgc1", precision))


def test_gcond_cuda_vs_oracle_inprocess():
    """Same seeded inputs through the oracle restatement and the CUDA path, without going through fixtures."""
    from graphslim_b200 import data as gdata
    from graphslim_b200.reduction import create_reducer
    from oracle import gcond_oracle as G
    name = "mini_sgc2_arxiv"
    args = helpers.case_args(name, device="cuda", save_init=False, progress=False, gemm_precision=0)
    args.epochs = 1
    raw = helpers.case_graph(name)
    helpers.seed_everything(args.seed)
    orc = G.GCondOracle(G.prepare_data(raw, args.dataset, args.pre_norm), args)
    ref_losses = orc.reduce()
    got = []
    helpers.seed_everything(args.seed)
    data = gdata.TransAndInd(raw, args.dataset, args.pre_norm)
    agent = create_reducer(args.method, setting=args.setting, data=data, args=args)
    agent.trace = lambda kind, **kw: got.append(float(kw["loss"].item())) if kind == "grads" else None
    agent.reduce(data, verbose=False)
    np.testing.assert_allclose(got[:2], ref_losses[:2], rtol=1e-4)
    np.testing.assert_allclose(got, ref_losses, rtol=1e-2)
    adj, feat, lab = orc.result()
    assert torch.equal(data.labels_syn.cpu(), lab)
    assert data.adj_syn.shape == adj.shape and data.feat_syn.shape == feat.shape


def test_single_bf16_mode_stated_bound():
    """gemm_precision 2 (single BF16 product on tcgen05): the explicitly looser mode, 3e-2."""
    check_against_golden(*run_case("mini_sgc2_arxiv", epochs=1, device="cuda", gemm_precision=2),
                         first_tol=3e-2, traj_tol=1e-1)


@pytest.mark.parametrize("name", ["mini_sgc1_trans", "mini_sgc2_arxiv", "mini_gcn_flickr", "mini_gcondx_mse"])
def test_inner_loop_cuda_graph_matches_stepwise(name):
    """The inner loop replayed from a captured CUDA graph (device-table Adam) runs the same kernels in the same order
    with the same scalars as the step-by-step path.  Where two step-by-step runs are bit-identical (no atomically
    accumulated product on the path) the graph run must be bit-identical too; otherwise it must agree as closely as
    they do.  The graph must actually have been used."""
    from graphslim_b200 import data as gdata
    from graphslim_b200.reduction import create_reducer

    def run(graphs):
        # grouped_mn=False: the deterministic per-class products, so that "two step-by-step runs" is a meaningful
        # yardstick (the TMA-fed grouped kernel combines CTAs with float atomics; Adam amplifies that last-bit noise to
        # ~1e-2 of the adjacency within two epochs, graph or no graph -- benchmarks/graph_vs_step_probe.py)
        args = helpers.case_args(name, device="cuda", save_init=False, progress=False, gemm_precision=1,
                                 cuda_graphs=graphs, grouped_mn=False, inner_chain=False)
        args.epochs = 2
        raw = helpers.case_graph(name)
        helpers.seed_everything(args.seed)
        data = gdata.TransAndInd(raw, args.dataset, args.pre_norm)
        agent = create_reducer(args.method, setting=args.setting, data=data, args=args)
        agent.reduce(data, verbose=False)
        torch.cuda.synchronize()
        if graphs:
            assert agent.inner.graph is not None, getattr(agent.inner, "capture_error", "graph not captured")
            assert agent.inner.replays > 0
            assert agent.match_graph.graph is not None, getattr(agent.match_graph, "capture_error", "not captured")
            assert agent.match_graph.replays > 0
        else:
            assert agent.inner.graph is None and agent.match_graph.graph is None
        return [data.feat_syn.cpu().clone(), data.adj_syn.cpu().clone()] + [w.cpu().clone() for w in agent.inner.W]

    a, b, g = run(False), run(False), run(True)
    if all(torch.equal(x, y) for x, y in zip(a, b)):
        for x, y in zip(a, g):
            assert torch.equal(x, y)
    else:
        for x, y, z in zip(a, b, g):
            # atomically accumulated products make two identical runs differ in the last bits, and Adam amplifies that
            # over the epochs: the graph run only has to stay inside the same (generously scaled) envelope
            noise = float((x - y).abs().max())
            scale = float(x.abs().max()) + 1e-30
            assert float((x - z).abs().max()) <= max(50 * noise, 2e-3 * scale)
