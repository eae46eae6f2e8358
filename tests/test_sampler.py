"""Product host sampler (csrc/host_sampler.cpp) against the oracle restatement of torch_sparse sample_adj, plus the
class-sharding property: a rank that materialises only some classes sees exactly the blocks of a full run and
leaves both RNG streams in the same state."""
import numpy as np
import pytest
import torch

from graphslim_b200 import synth
from graphslim_b200.sampler import ClassSampler
from oracle import gcond_oracle as G
from oracle import hostlib
from tests.test_engine_emulated import split_batch


def _graph(seed, n=3000, e=20000):
    raw = synth.make_graph(n=n, und_edges=e, d=4, c=6, split=(1500, 500, 1000), seed=seed)
    data = G.prepare_data(raw, "cora", False)
    adj = G.normalize_sparse(data.adj_full)
    lt = data.labels_train.numpy()
    members = [data.idx_train.numpy()[lt == c] for c in range(6)]
    return raw, data, adj, members


@pytest.mark.parametrize("dataset,nlayers", [("cora", 2), ("flickr", 2), ("cora", 3), ("cora", 1)])
def test_sampler_matches_oracle_bit_exact(dataset, nlayers):
    raw, data, adj, members = _graph(1)
    s = ClassSampler(adj.rowptr, adj.col.astype(np.int32), adj.val, members, dataset, nlayers, "cpu")
    s.set_labels(raw.y.numpy().astype(np.int32))
    for trial in range(3):
        np.random.seed(10 + trial)
        torch.manual_seed(20 + trial)
        rb = s.sample()
        got = split_batch(rb, 6)
        end_np, end_t = np.random.randint(1 << 30), int(torch.randint(0, 1 << 30, (1,)))
        np.random.seed(10 + trial)
        torch.manual_seed(20 + trial)
        sizes = G.fanouts(dataset, nlayers)
        for c in range(6):
            batch = np.random.permutation(members[c])[:256].astype(np.int64)
            n_id, blocks = batch, []
            for k in sizes:
                rp, col, new_ids, e_id = hostlib.sample_adj(adj.rowptr, adj.col, n_id, k)
                blocks.append((rp, col, adj.val[e_id]))
                n_id = new_ids
            bs, nid_got, blocks_got = got[c]
            assert bs == batch.size
            assert np.array_equal(nid_got, n_id)
            for (a, b, v), (a2, b2, v2) in zip(blocks[::-1], blocks_got):
                assert np.array_equal(a, a2) and np.array_equal(b, b2) and np.array_equal(v, v2)
            lab = rb.labels.numpy()[rb.seg[0][c]:rb.seg[0][c] + rb.cnt[0][c]]
            assert np.array_equal(lab, raw.y.numpy()[batch])
        assert end_np == np.random.randint(1 << 30) and end_t == int(torch.randint(0, 1 << 30, (1,)))


def test_transposed_blocks_are_transposes():
    raw, data, adj, members = _graph(2)
    s = ClassSampler(adj.rowptr, adj.col.astype(np.int32), adj.val, members, "cora", 2, "cpu")
    np.random.seed(0)
    torch.manual_seed(0)
    rb = s.sample()
    import scipy.sparse as sp
    for blk in rb.blocks:
        a = sp.csr_matrix((blk.csr.val.numpy(), blk.csr.col.numpy(), blk.csr.rowptr.numpy()),
                          shape=(blk.csr.n_rows, blk.csr.n_cols))
        t = sp.csr_matrix((blk.csr_t.val.numpy(), blk.csr_t.col.numpy(), blk.csr_t.rowptr.numpy()),
                          shape=(blk.csr_t.n_rows, blk.csr_t.n_cols))
        assert (a.T != t).nnz == 0
    outer = rb.blocks_fwd[0]
    g = outer.with_global_cols()
    assert np.array_equal(g.col.numpy(), rb.nid.numpy()[outer.csr.col.numpy()])


def test_class_sharding_replays_the_stream():
    raw, data, adj, members = _graph(3)
    s = ClassSampler(adj.rowptr, adj.col.astype(np.int32), adj.val, members, "cora", 2, "cpu")
    np.random.seed(5)
    torch.manual_seed(6)
    full = split_batch(s.sample(), 6)
    end = (np.random.randint(1 << 30), int(torch.randint(0, 1 << 30, (1,))))
    for mask in ([1, 0, 1, 0, 0, 1], [0, 1, 0, 1, 1, 0], [0, 0, 0, 0, 0, 1]):
        np.random.seed(5)
        torch.manual_seed(6)
        rb = s.sample(np.array(mask, dtype=np.uint8))
        assert rb.class_ids.tolist() == [c for c in range(6) if mask[c]]
        # rebuild a full-width seg so split_batch can index by class id
        part = split_batch_subset(rb, mask)
        for c in range(6):
            if mask[c]:
                bs, nid, blocks = part[c]
                assert bs == full[c][0] and np.array_equal(nid, full[c][1])
                for x, y in zip(blocks, full[c][2]):
                    assert all(np.array_equal(p, q) for p, q in zip(x, y))
        assert end == (np.random.randint(1 << 30), int(torch.randint(0, 1 << 30, (1,))))


def split_batch_subset(rb, mask):
    keep = [c for c in range(len(mask)) if mask[c]]
    pieces = split_batch(rb, len(keep))
    return {c: pieces[i] for i, c in enumerate(keep)}


def test_empty_and_tiny_classes():
    raw, data, adj, members = _graph(4)
    members = [m[:1] if i == 0 else m for i, m in enumerate(members)]     # a one-node class
    s = ClassSampler(adj.rowptr, adj.col.astype(np.int32), adj.val, members, "cora", 2, "cpu")
    np.random.seed(0)
    torch.manual_seed(0)
    rb = s.sample()
    assert int(rb.cnt[0][0]) == 1 and int(rb.seg[0][1]) == 64
    assert float(rb.inv_b[0]) == 1.0 and float(rb.inv_b[1:64].abs().max()) == 0.0   # pad targets carry no weight
