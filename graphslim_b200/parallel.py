"""Class-sharded GCond across the GPUs of one box (one process per GPU, torch.distributed).

The matching loss is a sum of independent per-class terms (graphslim/condensation/gcond_base.py:210-239), so
classes are dealt to ranks; every rank keeps a full replica of feat_syn / PGE / the condense model (same seed,
same RNG streams) and computes d loss / d feat_syn and d loss / d A_hat for its classes only.  One all-reduce per
outer step sums those partials (N'xd + N'xN' + 1 floats: 1-4 MB, latency-bound on NVLink).  The adjacency generator
(PGE) is sharded as well: its N'^2 pair rows are dealt to the ranks by slices of the first index, BatchNorm statistics
and the linear backward reductions travel as small all-gathers / all-reduces, and the adjacency rows are all-gathered
(pge.PGE.enable_row_sharding).  The optimiser steps and the inner loop run replicated and stay bit-identical across
ranks.

Index selection stays bit exact under sharding: every rank draws every class batch and replays the neighbour
sampler's random stream, materialising only its own classes (csrc/host_sampler.cpp).
"""
import numpy as np
import torch
import torch.distributed as dist

from .condensation.doscond import DosCond
from .condensation.doscondx import DosCondX
from .condensation.gcond import GCond
from .condensation.gcondx import GCondX


def partition_classes(class_sizes, world, batch=256):
    """Greedy longest-processing-time deal of classes to ranks, cost ~ sampled targets min(|class|, 256)."""
    cost = np.minimum(np.asarray(class_sizes, dtype=np.int64), batch)
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    owner = np.zeros(len(cost), dtype=np.int64)
    for c in order:
        r = int(np.argmin(load))
        owner[c] = r
        load[r] += cost[c]
    return [sorted(int(c) for c in np.nonzero(owner == r)[0]) for r in range(world)]


class _Sharded:
    def _init_shard(self, data, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        lt = np.asarray(data.labels_train)
        sizes = [int((lt == c).sum()) for c in range(data.nclass)]
        self.class_partition = partition_classes(sizes, self.world)
        self.owned_classes = self.class_partition[self.rank]
        if not self.owned_classes:
            raise ValueError(f"rank {self.rank} owns no class: world size {self.world} exceeds nclass {data.nclass}")
        # the adjacency generator is the largest replicated piece: deal its N'^2 pair rows to the ranks too
        self.pge_sharded = False
        if getattr(self.args, "shard_pge", True) and not self.x_variant and getattr(self, "pge", None) is not None:
            self.pge_sharded = self.pge.enable_row_sharding(group)

    def sync_replicas(self):
        """Rank 0's copy of everything the ranks update in lock step (synthetic features, PGE parameters, both Adam
        states) is broadcast to the others.  The replicas compute the same updates from the same all-reduced
        gradients, but products that accumulate split-K partials with atomics differ in their last bits from rank to
        rank; one broadcast per epoch (a few MB) keeps that from compounding through Adam over hundreds of epochs."""
        tensors = [self.feat_syn] + list(self.optimizer_feat.m) + list(self.optimizer_feat.v)
        if not self.x_variant:
            tensors += list(self.pge.parameters()) + list(self.optimizer_pge.m) + list(self.optimizer_pge.v)
        flat = torch.cat([t.reshape(-1) for t in tensors])
        dist.broadcast(flat, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        off = 0
        for t in tensors:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()

    def run_epoch(self, it):
        super().run_epoch(it)
        if self.world > 1:
            self.sync_replicas()

    def reduce_partials(self, loss, dX, dA):
        parts = [loss.reshape(-1), dX.reshape(-1)] + ([dA.reshape(-1)] if dA is not None else [])
        flat = torch.cat(parts)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        n0, n1 = loss.numel(), loss.numel() + dX.numel()
        loss_r = flat[:n0].view_as(loss)
        dX_r = flat[n0:n1].view_as(dX)
        dA_r = flat[n1:].view_as(dA) if dA is not None else None
        self.allreduce_bytes = flat.numel() * 4
        return loss_r, dX_r, dA_r


class ShardedGCond(_Sharded, GCond):
    def __init__(self, setting, data, args, group=None, **kwargs):
        GCond.__init__(self, setting, data, args, **kwargs)
        self._init_shard(data, group)


class ShardedGCondX(_Sharded, GCondX):
    def __init__(self, setting, data, args, group=None, **kwargs):
        GCondX.__init__(self, setting, data, args, **kwargs)
        self._init_shard(data, group)


class ShardedDosCond(_Sharded, DosCond):
    def __init__(self, setting, data, args, group=None, **kwargs):
        DosCond.__init__(self, setting, data, args, **kwargs)
        self._init_shard(data, group)


class ShardedDosCondX(_Sharded, DosCondX):
    def __init__(self, setting, data, args, group=None, **kwargs):
        DosCondX.__init__(self, setting, data, args, **kwargs)
        self._init_shard(data, group)


SHARDED = {"gcond": ShardedGCond, "gcondx": ShardedGCondX, "doscond": ShardedDosCond, "doscondx": ShardedDosCondX}


# ============================================================================================================
# Row-partitioned full-graph propagation  Y = A_hat X  
# ============================================================================================================
def partition_rows_by_nnz(rowptr, world):
    """Contiguous row blocks of (nearly) equal work: bounds[r]..bounds[r+1] are rank r's rows.  Work of a row is its
    non-zero count plus one (the row's own load/store), so empty rows still spread out."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    n = rowptr.size - 1
    cum = rowptr[1:] + np.arange(1, n + 1, dtype=np.int64)            # work of rows 0..i inclusive
    total = int(cum[-1]) if n else 0
    bounds = [0]
    for r in range(1, world):
        b = int(np.searchsorted(cum, (total * r + world - 1) // world, side="left")) + 1 if n else 0
        bounds.append(min(max(b, bounds[-1]), n))
    bounds.append(n)
    return np.asarray(bounds, dtype=np.int64)


def feature_slabs(F, n_slabs, quantum=4):
    """Column ranges [(c0, c1)] of roughly equal width whose starts are multiples of `quantum` floats (16 B:
    the float4 alignment of the SpMM kernel)."""
    units = -(-F // quantum)
    n_slabs = max(1, min(int(n_slabs), units))
    cuts = [min(F, (units * s // n_slabs) * quantum) for s in range(n_slabs)] + [F]
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]


class RowPartitionedSpmm:
    """Y = A_hat X with the rows of A_hat (CSR) dealt to the ranks in contiguous nnz-balanced blocks and X sharded by
    the same row blocks.  forward: all-gather of the X shards (the halo is every remote row: NVSwitch is uniform),
    then the local warp-per-row kernel.  The gather is issued in feature-column slabs so that the transfer of slab
    s+1 (NCCL, on its own stream) overlaps the SpMM of slab s; each output row is still produced by one rank in CSR
    order, so the result is bit-identical to the single-GPU kernel.  backward (d/dX): the transposed local block
    applied to dY_local (no atomics), then a reduce-scatter of the partial dX.

    Index space: every shard is padded to `pad` = max shard rows, global row g of rank r sits at r*pad + (g -
    bounds[r]); column ids of the local CSR are remapped once on the host, so collectives work on equal shards.

    `spmm(csr, X, out)` is the local kernel (CudaOps.spmm on a GPU; tests on CPU inject the oracle's).
    """

    def __init__(self, rowptr, col, val, *, rank, world, group=None, device="cuda", make_csr=None, spmm=None,
                 n_slabs=4, long_row_nnz=64):
        import scipy.sparse as sp
        rowptr = np.asarray(rowptr, dtype=np.int64)
        col = np.asarray(col, dtype=np.int64)
        val = np.asarray(val, dtype=np.float32)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.device = torch.device(device)
        self.n = n = rowptr.size - 1
        self.bounds = b = partition_rows_by_nnz(rowptr, world)
        self.lo, self.hi = int(b[rank]), int(b[rank + 1])
        self.rows_local = self.hi - self.lo
        self.pad = int(max(1, (b[1:] - b[:-1]).max()))
        self.n_slabs = int(n_slabs)
        e0, e1 = int(rowptr[self.lo]), int(rowptr[self.hi])
        lcol = col[e0:e1]
        owner = np.searchsorted(b, lcol, side="right") - 1
        pcol = owner * self.pad + (lcol - b[owner])                                  # padded index space
        lptr = rowptr[self.lo:self.hi + 1] - e0
        self.nnz_local = e1 - e0
        local = sp.csr_matrix((val[e0:e1], pcol, lptr), shape=(self.rows_local, world * self.pad))
        local_t = local.T.tocsr()
        local_t.sort_indices()
        self._make_csr = make_csr or self._device_csr
        self.long_row_nnz = long_row_nnz
        self.csr = self._make_csr(local)
        self.csr_t = self._make_csr(local_t)
        self._spmm = spmm
        self._bufs = {}

    # -- plumbing ------------------------------------------------------------------------------------------
    def _device_csr(self, m):
        from .graph_utils import build_row_chunks, chunks_to_device
        from .ops import Csr
        dev = self.device
        chunks = chunks_to_device(build_row_chunks(m.indptr, self.long_row_nnz), dev)
        return Csr(torch.from_numpy(m.indptr.astype(np.int32)).to(dev), torch.from_numpy(m.indices.astype(np.int32)).to(dev),
                   torch.from_numpy(m.data.astype(np.float32)).to(dev), m.shape[0], m.shape[1], chunks)

    def _buf(self, key, shape):
        t = self._bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.zeros(*shape, dtype=torch.float32, device=self.device)
            self._bufs[key] = t
        return t

    def shard(self, X_full):
        """Rows of a replicated matrix that this rank owns."""
        return X_full[self.lo:self.hi]

    # -- forward -------------------------------------------------------------------------------------------
    def forward(self, X_local, out=None):
        if X_local.shape[0] != self.rows_local:
            raise ValueError(f"X_local has {X_local.shape[0]} rows, this rank owns {self.rows_local}")
        F = X_local.shape[1]
        Y = out if out is not None else torch.empty(self.rows_local, F, dtype=torch.float32, device=self.device)
        slabs = feature_slabs(F, self.n_slabs)
        works = [None] * len(slabs)

        def issue(s):
            c0, c1 = slabs[s]
            send = self._buf(("send", s, c1 - c0), (self.pad, c1 - c0))
            send[:self.rows_local].copy_(X_local[:, c0:c1])
            full = self._buf(("full", s, c1 - c0), (self.world * self.pad, c1 - c0))
            works[s] = (dist.all_gather_into_tensor(full, send, group=self.group, async_op=True), full)

        issue(0)
        for s, (c0, c1) in enumerate(slabs):
            if s + 1 < len(slabs):
                issue(s + 1)                       # transfer of the next slab overlaps this slab's SpMM
            work, full = works[s]
            work.wait()                            # stream-ordered for NCCL (the host does not block)
            self._spmm(self.csr, full, Y[:, c0:c1])
        return Y

    # -- backward w.r.t. X ---------------------------------------------------------------------------------
    def backward(self, dY_local):
        """Rows lo..hi of A_hat^T dY, dY row-sharded like Y."""
        F = dY_local.shape[1]
        part = self._buf(("part", F), (self.world * self.pad, F))
        self._spmm(self.csr_t, dY_local.contiguous(), part)
        mine = self._buf(("mine", F), (self.pad, F))
        dist.reduce_scatter_tensor(mine, part, op=dist.ReduceOp.SUM, group=self.group)
        return mine[:self.rows_local].clone()

    # -- accounting (bench) --------------------------------------------------------------------------------
    def bytes_model(self, F):
        """(algorithmic HBM bytes of the local kernel, bytes received over NVLink) per forward on this rank."""
        alg = 4 * (self.rows_local + 1) + 8 * self.nnz_local + 4 * F * self.world * self.pad + 4 * F * self.rows_local
        recv = 4 * F * self.pad * (self.world - 1)
        return alg, recv
