"""Device sampler (csrc/device_sampler.cu) against the host sampler (csrc/host_sampler.cpp, itself pinned bit-exact to
the oracle restatement of torch_sparse sample_adj in tests/test_sampler.py): every packed array of every step must be
identical, and both random streams must end in the same state -- also under class sharding and when prefetched."""
import numpy as np
import pytest
import torch

from graphslim_b200 import synth
from graphslim_b200.ops import Csr
from graphslim_b200.sampler import ClassSampler, DeviceClassSampler
from oracle import gcond_oracle as G

pytestmark = pytest.mark.gpu


def _graph(seed, n=6000, e=60000, c=6, hub=True):
    raw = synth.make_graph(n=n, und_edges=e, d=4, c=c, split=(n // 2, n // 6, n // 3), seed=seed)
    data = G.prepare_data(raw, "cora", False)
    adj = G.normalize_sparse(data.adj_full)
    lt = data.labels_train.numpy()
    members = [data.idx_train.numpy()[lt == k] for k in range(c)]
    return raw, adj, members


def _pair(seed, dataset, nlayers, **kw):
    raw, adj, members = _graph(seed, **kw)
    lab = raw.y.numpy().astype(np.int32)
    host = ClassSampler(adj.rowptr, adj.col.astype(np.int32), adj.val, members, dataset, nlayers, "cpu")
    host.set_labels(lab)
    mk = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).cuda()
    csr = Csr(mk(adj.rowptr, np.int32), mk(adj.col, np.int32), mk(adj.val, np.float32), adj.rowptr.size - 1,
              adj.rowptr.size - 1)
    dev = DeviceClassSampler(csr, members, dataset, nlayers, "cuda", labels=torch.from_numpy(lab))
    return host, dev


def _same(a, b, nh):
    assert a.counts == b.counts
    eq = lambda x, y: np.array_equal(x.cpu().numpy(), y.cpu().numpy())
    for name in ("nid", "tcls", "inv_b", "target_ids", "labels", "cnt"):
        assert eq(getattr(a, name), getattr(b, name)), name
    for l in range(nh + 1):
        assert eq(a.seg[l], b.seg[l]), f"seg[{l}]"
    for h in range(nh):
        x, y = a.blocks[h], b.blocks[h]
        for part in ("rowptr", "col", "val"):
            assert eq(getattr(x.csr, part), getattr(y.csr, part)), f"hop {h} csr.{part}"
            assert eq(getattr(x.csr_t, part), getattr(y.csr_t, part)), f"hop {h} csr_t.{part}"
        assert (x._gcol is None) == (y._gcol is None)
        if x._gcol is not None:
            assert eq(x._gcol, y._gcol), f"hop {h} gcol"


@pytest.mark.parametrize("dataset,nlayers", [("cora", 2), ("flickr", 2), ("cora", 3), ("cora", 1)])
@pytest.mark.parametrize("mask", [None, [1, 0, 1, 1, 0, 0]])
def test_device_sampler_is_bit_identical_to_host_sampler(dataset, nlayers, mask):
    host, dev = _pair(3, dataset, nlayers)
    for trial in range(3):                       # consecutive steps: the generator position carries over
        np.random.seed(100 + trial)
        torch.manual_seed(200 + trial)
        torch.randint(0, 10, (trial * 211 + 5,))     # start mid-block, different offsets
        a = host.sample(mask)
        end_h = (np.random.randint(1 << 30), int(torch.randint(0, 1 << 30, (1,))))
        np.random.seed(100 + trial)
        torch.manual_seed(200 + trial)
        torch.randint(0, 10, (trial * 211 + 5,))
        b = dev.sample(mask)
        end_d = (np.random.randint(1 << 30), int(torch.randint(0, 1 << 30, (1,))))
        _same(a, b, host.nh)
        assert end_h == end_d


def test_device_sampler_prefetch_matches_host_loop_and_dense_hubs():
    # denser graph (most rows exceed the fan-out, a few do not) and small classes (< 256 members)
    host, dev = _pair(5, "flickr", 2, n=3000, e=150000, c=7)
    steps = 6
    np.random.seed(7)
    torch.manual_seed(8)
    ref = [host.sample() for _ in range(steps)]
    end_h = (np.random.randint(1 << 30), int(torch.randint(0, 1 << 30, (1,))))
    np.random.seed(7)
    torch.manual_seed(8)
    pf = dev.prefetch(steps)
    for i in range(steps):
        rb = pf.next()
        _same(ref[i], rb, host.nh)
    pf.join()
    end_d = (np.random.randint(1 << 30), int(torch.randint(0, 1 << 30, (1,))))
    assert end_h == end_d
