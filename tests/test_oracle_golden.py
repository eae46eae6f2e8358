"""The oracle restatement (oracle/gcond_oracle.py) against fixtures produced by the UNMODIFIED
reference (oracle/make_goldens.py).  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import gcond_oracle as G
from oracle.cases import CASES
from tests import helpers


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return np.frombuffer(h.digest()[:8], dtype=np.uint64)[0]


FAST = [c for c in CASES if c != "cora_sgc1"]


@pytest.mark.parametrize("name", FAST + ["cora_sgc1"])
def test_oracle_matches_reference_golden(name):
    gold = helpers.golden(name)
    sub = int(gold["grad_subsample"])
    args = helpers.case_args(name)
    raw = helpers.case_graph(name)
    data = G.prepare_data(raw, args.dataset, args.pre_norm)
    seen = dict(digests=[], grads={}, model_init=[], samples={})

    def obs(kind, *a):
        if kind == "norm":
            seen["adj"] = a[0]
        elif kind == "sample":
            step, c, bs, n_id, blocks = a
            parts = [n_id]
            for b in blocks:
                parts += [b.rowptr, b.col, b.val]
            seen["digests"].append(_digest(*parts))
            if step == 0:
                seen["samples"][c] = (bs, n_id, blocks)
        elif kind == "grads":
            step, fg, pg = a
            if f"g{step}_feat" in gold:
                flat = np.concatenate([(g.numpy().ravel() if g is not None else np.zeros(0, np.float32)) for g in pg])
                seen["grads"][step] = (fg.numpy().copy()[:, ::sub], flat)
        elif kind == "model_init":
            seen["model_init"].append(a[1][::sub])

    helpers.seed_everything(args.seed)
    orc = G.GCondOracle(data, args, observer=obs)
    # integer / index work: bit exact
    assert np.array_equal(orc.labels_syn_np, gold["labels_syn"])
    assert list(orc.alloc.keys()) == gold["class_order"].tolist()
    assert list(orc.alloc.values()) == gold["class_count"].tolist()
    pge_init = np.concatenate([p.detach().numpy().ravel() for p in orc.pge.parameters()])[::sub]
    assert np.array_equal(pge_init, gold["pge_init"])
    losses = orc.reduce()
    adj = seen["adj"]
    assert np.array_equal(adj.rowptr, gold["adj_rowptr"])
    assert np.array_equal(adj.col, gold["adj_col"])
    assert np.array_equal(adj.val, gold["adj_val"])          # fp32 values bit exact
    assert np.array_equal(np.array(seen["digests"], dtype=np.uint64), gold["sample_digest"])
    for c, (bs, n_id, blocks) in seen["samples"].items():
        assert bs == int(gold[f"s0_c{c}_bs"])
        assert np.array_equal(n_id, gold[f"s0_c{c}_nid"])
        for h, b in enumerate(blocks):
            assert np.array_equal(b.rowptr, gold[f"s0_c{c}_h{h}_rowptr"])
            assert np.array_equal(b.col, gold[f"s0_c{c}_h{h}_col"])
            assert np.array_equal(b.val, gold[f"s0_c{c}_h{h}_val"])
    assert np.array_equal(np.stack(seen["model_init"]), gold["model_init"])
    # floating point: same torch build, same op order -> tight
    # The first steps start from identical state: tight.  Later steps sit on an Adam trajectory whose g/sqrt(v) feeds the
    # run-to-run noise of multi-threaded CPU reductions back into the parameters (the unmodified reference shows the
    # same spread between two runs of itself: up to 7e-5 on step 40 of this 48-step case), so the tail gets 3e-4.
    np.testing.assert_allclose(np.array(losses)[:4], gold["losses"][:4], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(np.array(losses), gold["losses"], rtol=3e-4, atol=1e-7)
    for step, (fg, pg) in seen["grads"].items():
        np.testing.assert_allclose(fg, gold[f"g{step}_feat"], rtol=1e-4, atol=1e-7 * (1 + np.abs(gold[f"g{step}_feat"]).max()))
        ref = gold[f"g{step}_pge"]
        # (the reference itself is not run-to-run deterministic below ~1e-5 of the gradient scale: two generator runs
        # of the mse/GCN DosCond case differ in the 8th digit of the second loss)
        np.testing.assert_allclose(pg[::sub], ref, rtol=1e-4, atol=2e-5 * (1e-30 + np.abs(ref).max()))
    feat_final = orc.feat_syn.detach().numpy()[:, ::sub]
    pge_final = np.concatenate([p.detach().numpy().ravel() for p in orc.pge.parameters()])[::sub]
    if CASES[name]["method"].startswith("doscond"):
        # features move from the very first step here (both optimisers step every time): Adam turns the reference's
        # own run-to-run noise on near-zero gradient entries into +-lr moves, so the end state is compared in norm
        rel = np.linalg.norm(feat_final - gold["feat_final"]) / np.linalg.norm(gold["feat_final"])
        assert rel < 5e-3, rel
        relp = np.linalg.norm(pge_final - gold["pge_final"]) / np.linalg.norm(gold["pge_final"])
        assert relp < 5e-3, relp
    else:
        # multi-threaded CPU reductions make the oracle (and the reference) bimodal run to run: mini_sgc1_trans ends
        # either 1.6e-7 from the fixture or, when one reduction splits differently, 3.2e-4 in relative Frobenius norm
        # with 111 of 1600 entries moved by up to 12 % (near-zero gradient entries, +-lr through Adam) while every loss
        # stays within 7e-5 -- so the end state is compared in norm here too
        rel = np.linalg.norm(feat_final - gold["feat_final"]) / np.linalg.norm(gold["feat_final"])
        assert rel < 2e-3, rel
        np.testing.assert_allclose(pge_final, gold["pge_final"], rtol=1e-3, atol=1e-5)
    # total RNG consumption identical
    assert np.array_equal(np.random.randint(0, 2**31 - 1, size=4).astype(np.int64), gold["np_rng_probe"])
    assert np.array_equal(torch.randint(0, 2**31 - 1, (4,)).numpy(), gold["torch_rng_probe"])
