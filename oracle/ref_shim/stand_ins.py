"""ORACLE (test infrastructure) -- CPU stand-ins for the third-party symbols the
reference's GCond path executes (see oracle/ref_shim/__init__.py).

Pinned upstream versions (requirements.lock:169-172 of the reference):
torch_sparse 0.6.18, torch_geometric 2.7.0.  Behaviour restated from their
published sources; torch_sparse itself is absent here, so this boundary is
"parity unpinned" (SURVEY.md section 8c).
"""
import numpy as np
import torch

from .. import hostlib


# --------------------------------------------------------------------------- torch_sparse
class _Storage:
    def __init__(self, owner):
        self._o = owner

    def value(self):
        return self._o._value

    def rowptr(self):
        return self._o._rowptr

    def col(self):
        return self._o._col

    def row(self):
        return self._o._row()


class SparseTensor:
    """Minimal CSR-backed stand-in for torch_sparse.SparseTensor (row-major sorted)."""

    def __init__(self, row=None, rowptr=None, col=None, value=None, sparse_sizes=None, is_sorted=False,
                 trust_data=False):
        assert col is not None
        col = col.long()
        if sparse_sizes is None:
            m = int(row.max()) + 1 if rowptr is None else rowptr.numel() - 1
            n = int(col.max()) + 1
            sparse_sizes = (m, n)
        self._sizes = (int(sparse_sizes[0]), int(sparse_sizes[1]))
        if rowptr is None:
            row = row.long()
            if not is_sorted:
                key = row * self._sizes[1] + col
                perm = torch.argsort(key, stable=True)
                row, col = row[perm], col[perm]
                value = value[perm] if value is not None else None
            counts = torch.bincount(row, minlength=self._sizes[0])
            rowptr = torch.zeros(self._sizes[0] + 1, dtype=torch.long)
            rowptr[1:] = torch.cumsum(counts, 0)
        self._rowptr = rowptr.long()
        self._col = col
        self._value = value
        self._t_cache = None

    # ---- construction helpers
    @classmethod
    def from_edge_index(cls, edge_index, edge_attr=None, sparse_sizes=None, is_sorted=False, trust_data=False):
        return cls(row=edge_index[0], col=edge_index[1], value=edge_attr, sparse_sizes=sparse_sizes,
                   is_sorted=is_sorted)

    # ---- accessors
    @property
    def storage(self):
        return _Storage(self)

    def _row(self):
        counts = self._rowptr[1:] - self._rowptr[:-1]
        return torch.repeat_interleave(torch.arange(self._sizes[0]), counts)

    def csr(self):
        return self._rowptr, self._col, self._value

    def coo(self):
        return self._row(), self._col, self._value

    def sparse_sizes(self):
        return self._sizes

    def sizes(self):
        return list(self._sizes)

    def size(self, dim=None):
        return self._sizes[dim] if dim is not None else self._sizes

    def nnz(self):
        return self._col.numel()

    def has_value(self):
        return self._value is not None

    def set_value(self, value, layout=None):
        return SparseTensor(rowptr=self._rowptr, col=self._col, value=value, sparse_sizes=self._sizes, is_sorted=True)

    @property
    def device(self):
        return self._col.device

    def to(self, *a, **k):
        return self

    def cpu(self):
        return self

    def t(self):
        if self._t_cache is None:
            row = self._row()
            self._t_cache = SparseTensor(row=self._col, col=row, value=self._value,
                                         sparse_sizes=(self._sizes[1], self._sizes[0]))
        return self._t_cache

    def to_dense(self):
        out = torch.zeros(self._sizes, dtype=self._value.dtype if self._value is not None else torch.float32)
        v = self._value if self._value is not None else torch.ones(self.nnz())
        out.index_put_((self._row(), self._col), v, accumulate=True)
        return out

    # ---- ops on the GCond path
    def sample_adj(self, subset, num_neighbors, replace=False):
        assert not replace
        rp, c, n_id, e_id = hostlib.sample_adj(self._rowptr.numpy(), self._col.numpy(), subset.numpy(),
                                               int(num_neighbors))
        e_id = torch.from_numpy(e_id)
        value = self._value[e_id] if self._value is not None else None
        out = SparseTensor(rowptr=torch.from_numpy(rp), col=torch.from_numpy(c), value=value,
                           sparse_sizes=(subset.numel(), n_id.size), is_sorted=True)
        return out, torch.from_numpy(n_id)

    def __matmul__(self, other):
        return matmul(self, other)

    def matmul(self, other):
        return matmul(self, other)


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sp):
        ctx.sp = sp
        y = hostlib.spmm_csr(sp._rowptr.numpy(), sp._col.numpy(), sp._value.detach().numpy(), x.detach().numpy())
        return torch.from_numpy(y)

    @staticmethod
    def backward(ctx, g):
        return _SpMM.apply(g.contiguous(), ctx.sp.t()), None


def _spspmm(a, b):
    """torch_sparse.matmul(SparseTensor, SparseTensor) (csrc spspmm: fp32 row-by-row accumulation): scipy's CSR
    product in float32, columns sorted -- used by the `--agg` init only (model_free_coreset_base.py:20-24)."""
    import scipy.sparse as sp
    ma = sp.csr_matrix((a._value.numpy(), a._col.numpy(), a._rowptr.numpy()), shape=tuple(a._sizes))
    mb = sp.csr_matrix((b._value.numpy(), b._col.numpy(), b._rowptr.numpy()), shape=tuple(b._sizes))
    m = (ma @ mb).tocsr().astype(np.float32)
    m.sort_indices()
    return SparseTensor(rowptr=torch.from_numpy(m.indptr.astype(np.int64)), col=torch.from_numpy(m.indices.astype(np.int64)),
                        value=torch.from_numpy(m.data), sparse_sizes=m.shape, is_sorted=True)


def matmul(src, other, reduce="sum"):
    assert reduce == "sum"
    if isinstance(src, SparseTensor) and isinstance(other, SparseTensor):
        return _spspmm(src, other)
    assert isinstance(src, SparseTensor) and isinstance(other, torch.Tensor), "stand-in covers sparse @ dense / sparse"
    assert other.dim() == 2 and other.dtype == torch.float32
    return _SpMM.apply(other.contiguous(), src)


# --------------------------------------------------------------------------- torch_geometric
class NeighborSampler:
    """torch_geometric.loader.NeighborSampler restricted to what loader.py:212-223 uses
    (SparseTensor input, return_e_id=False, ``sample(batch)`` called directly)."""

    def __init__(self, edge_index, sizes, node_idx=None, num_nodes=None, return_e_id=True, transform=None, **kwargs):
        assert isinstance(edge_index, SparseTensor) and not return_e_id
        self.adj_t = edge_index
        self.sizes = list(sizes)
        self.node_idx = node_idx

    def sample(self, batch):
        if not isinstance(batch, torch.Tensor):
            batch = torch.tensor(batch)
        batch_size = len(batch)
        adjs = []
        n_id = batch
        for size in self.sizes:
            adj_t, n_id = self.adj_t.sample_adj(n_id, size, replace=False)
            e_id = adj_t.storage.value()
            shape = adj_t.sparse_sizes()[::-1]
            adjs.append((adj_t, e_id, shape))
        adjs = adjs[0] if len(adjs) == 1 else adjs[::-1]
        return batch_size, n_id, adjs


def to_undirected(edge_index, num_nodes=None, **kw):
    """Symmetrise + coalesce (sorted by row then col), as torch_geometric.utils.to_undirected."""
    if isinstance(num_nodes, torch.Tensor):
        num_nodes = int(num_nodes)
    row, col = edge_index[0], edge_index[1]
    r = torch.cat([row, col])
    c = torch.cat([col, row])
    n = int(max(r.max(), c.max())) + 1 if num_nodes is None else int(num_nodes)
    key = torch.unique(r * n + c, sorted=True)
    return torch.stack([key // n, key % n])


def populate(module):
    name = module.__name__
    if name == "torch_sparse":
        module.SparseTensor = SparseTensor
        module.matmul = matmul
    elif name == "torch_geometric.loader":
        module.NeighborSampler = NeighborSampler
    elif name == "torch_geometric.utils":
        module.to_undirected = to_undirected
