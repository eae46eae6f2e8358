"""Seeded synthetic graphs with the shapes of the datasets the GCond configs are quoted on.

There is no network in the build/bench environment, so the real Cora / ogbn-arxiv /
Flickr / Reddit files are unavailable; every benchmark and parity run uses graphs
from this generator: an undirected simple power-law
(Chung-Lu style) graph with a prescribed number of directed non-zeros, N(0,1)
features, skewed class sizes (every class present in train) and splits sized like
the real datasets.  The object returned mimics the PyG ``Data`` fields the
reference's ``TransAndInd`` consumes (graphslim/dataset/loader.py:100-135).
"""
from types import SimpleNamespace

import numpy as np
import torch

# name -> (nodes, undirected edges, feats, classes, (train, val, test), dataset flag used by the reference)
SHAPES = {
    "cora": dict(n=2708, und_edges=5278, d=1433, c=7, split=(140, 500, 1000), per_class_train=20),
    "ogbn-arxiv": dict(n=169343, und_edges=1166243, d=128, c=40, split=(90941, 29799, 48603)),
    "flickr": dict(n=89250, und_edges=449878, d=500, c=7, split=(44625, 22312, 22313)),
    "reddit": dict(n=232965, und_edges=57307946, d=602, c=41, split=(153932, 23699, 55334)),
}


def _powerlaw_edges_cuda(n, und_edges, exponent, seed):
    """Same construction with torch on the GPU (seconds instead of minutes at the Reddit scale).  Seeded and
    deterministic on a given device type, but a different stream than the numpy path."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(seed)
    w = torch.arange(1, n + 1, device=dev, dtype=torch.float64) ** (-1.0 / (exponent - 1.0))
    w = w[torch.randperm(n, device=dev, generator=g)]
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    keys = torch.empty(0, dtype=torch.int64, device=dev)
    need = int(und_edges)
    while keys.numel() < need:
        m = int((need - keys.numel()) * 1.3) + 64
        u = torch.searchsorted(cdf, torch.rand(m, device=dev, dtype=torch.float64, generator=g)).clamp_(max=n - 1)
        v = torch.searchsorted(cdf, torch.rand(m, device=dev, dtype=torch.float64, generator=g)).clamp_(max=n - 1)
        keep = u != v
        lo, hi = torch.minimum(u[keep], v[keep]), torch.maximum(u[keep], v[keep])
        keys = torch.unique(torch.cat([keys, lo * n + hi]))
    if keys.numel() > need:
        keys = keys[torch.randperm(keys.numel(), device=dev, generator=g)[:need]]
    lo, hi = (keys // n).cpu().numpy(), (keys % n).cpu().numpy()
    return np.stack([np.concatenate([lo, hi]), np.concatenate([hi, lo])])


def powerlaw_edges(n, und_edges, exponent=2.3, seed=0, allow_cuda=True):
    """Return a (2, 2*und_edges) int64 edge_index of a simple undirected power-law graph.

    Both directions are present, no self loops, no duplicates (edge order is unspecified: every consumer builds a
    CSR from it).  Graphs of more than 5M edges are generated on the GPU when one is present (benchmark shapes only;
    every parity fixture is far below that and always takes the numpy path).
    """
    if allow_cuda and und_edges > 5_000_000 and torch.cuda.is_available():
        return _powerlaw_edges_cuda(n, und_edges, exponent, seed)
    rng = np.random.default_rng(seed)
    w = np.arange(1, n + 1, dtype=np.float64) ** (-1.0 / (exponent - 1.0))
    rng.shuffle(w)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    keys = np.empty(0, dtype=np.int64)
    need = int(und_edges)
    max_possible = n * (n - 1) // 2
    if need > max_possible:
        raise ValueError("more edges requested than a simple graph on n nodes can hold")
    while keys.size < need:
        m = int((need - keys.size) * 1.3) + 64
        u = np.searchsorted(cdf, rng.random(m), side="right").astype(np.int64)
        v = np.searchsorted(cdf, rng.random(m), side="right").astype(np.int64)
        np.minimum(u, n - 1, out=u)
        np.minimum(v, n - 1, out=v)
        keep = u != v
        lo = np.minimum(u[keep], v[keep])
        hi = np.maximum(u[keep], v[keep])
        keys = np.unique(np.concatenate([keys, lo * n + hi]))
    if keys.size > need:
        keys = np.sort(rng.permutation(keys)[:need])
    lo, hi = keys // n, keys % n
    return np.stack([np.concatenate([lo, hi]), np.concatenate([hi, lo])])


def make_graph(name=None, *, n=None, und_edges=None, d=None, c=None, split=None, per_class_train=None,
               exponent=2.3, seed=0, class_skew=0.7):
    """Build a PyG-like namespace: x (n,d) fp32, y (n,) int64, edge_index (2,nnz) int64, masks, idx_*."""
    if name is not None:
        spec = dict(SHAPES[name])
        n = spec["n"] if n is None else n
        und_edges = spec["und_edges"] if und_edges is None else und_edges
        d = spec["d"] if d is None else d
        c = spec["c"] if c is None else c
        split = spec["split"] if split is None else split
        per_class_train = spec.get("per_class_train") if per_class_train is None else per_class_train
    rng = np.random.default_rng(seed + 1)
    ei = powerlaw_edges(n, und_edges, exponent=exponent, seed=seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    p = np.arange(1, c + 1, dtype=np.float64) ** (-class_skew)
    p = rng.permutation(p / p.sum())
    y = rng.choice(c, size=n, p=p).astype(np.int64)
    n_train, n_val, n_test = split
    perm = rng.permutation(n)
    if per_class_train:
        train = []
        for k in range(c):
            members = perm[y[perm] == k]
            if members.size < per_class_train:
                raise ValueError("class too small for the requested per-class train size")
            train.append(members[:per_class_train])
        train = np.concatenate(train)
        rest = perm[~np.isin(perm, train)]
    else:
        train = perm[:n_train]
        rest = perm[n_train:]
        # every class must occur in train (gcond_base.py:237 KeyError otherwise)
        for k in range(c):
            if not (y[train] == k).any():
                y[train[k]] = k
    val, test = rest[:n_val], rest[n_val:n_val + n_test]
    masks = []
    for idx in (train, val, test):
        m = np.zeros(n, dtype=bool)
        m[idx] = True
        masks.append(torch.from_numpy(m))
    data = SimpleNamespace()
    data.x = torch.from_numpy(x)
    data.y = torch.from_numpy(y)
    data.edge_index = torch.from_numpy(ei)
    data.num_nodes = n
    data.train_mask, data.val_mask, data.test_mask = masks
    data.idx_train = data.train_mask.nonzero().view(-1)
    data.idx_val = data.val_mask.nonzero().view(-1)
    data.idx_test = data.test_mask.nonzero().view(-1)
    data.num_classes = c
    return data
