"""ORACLE (test infrastructure, not product code) -- CPU restatement of the reference's
GCond / GCondX condensation loop in plain torch (CPU autograd) + numpy/scipy.

It is the checker for the CUDA path and the ``cpu_baseline`` / ``--impl reference`` arm of
bench.py.  Only tests/, ``__graft_entry__.smoke()`` and those bench legs may import it.

Pinning: tests/test_oracle_golden.py checks this restatement against fixtures produced by
the UNMODIFIED reference (oracle/make_goldens.py through oracle/ref_shim).  The neighbour
sampler underneath both is a restatement of torch_sparse 0.6.18 (absent from the
container): *parity unpinned* at that one boundary.

Reference lines restated (all under /root/reference/graphslim):
  data object ............ dataset/loader.py:100-135, dataset/convertor.py:71-75
  class sampler .......... dataset/loader.py:187-224
  label allocation ....... condensation/gcond_base.py:79-115
  random init ............ sparsification/random.py:9-17, model_free_coreset_base.py:16-61
  normalisation .......... utils.py:403-458 (sparse: float64 scipy; dense: two diag matmuls)
  SGC / GCN / layers ..... models/sgc.py:12-57, models/gcn.py:8-23, models/base.py:44-78,
                           models/layers.py:17-56,354-386
  PGE .................... models/parametrized_adj.py:7-86
  matching loss .......... condensation/utils.py:12-106
  per-class matching ..... condensation/gcond_base.py:156-241
  loops .................. condensation/gcond.py:17-81, condensation/gcondx.py:17-79, doscond.py:17-65, doscondx.py:19-63
"""
import math
from collections import Counter
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import hostlib


# ----------------------------------------------------------------------------- data object
def prepare_data(raw, dataset, pre_norm=True):
    """loader.py:100-135 for a PyG-like ``raw`` (x, y, edge_index, num_nodes, idx_*)."""
    d = SimpleNamespace()
    ei = raw.edge_index.numpy()
    n = int(raw.num_nodes)
    if dataset in ("flickr", "reddit", "ogbn-arxiv"):
        # loader.py:113-119 symmetrises edge_index only (adj_full was already built) and standardises
        mu = raw.x[raw.idx_train].numpy().astype(np.float64)
        mean = mu.mean(0)
        var = mu.var(0)
        scale = np.sqrt(var)
        scale[scale == 0.0] = 1.0
        feat = ((raw.x.numpy() - mean) / scale)
        feat_full = torch.from_numpy(feat).float()
    else:
        feat_full = raw.x
    if pre_norm and dataset in ("cora", "citeseer", "pubmed"):
        feat_full = F.normalize(feat_full, p=1, dim=1)
    adj_full = sp.coo_matrix((np.ones_like(ei[0]), (ei[0], ei[1])), shape=(n, n)).tocsr()  # convertor.py:71-75
    d.adj_full, d.feat_full, d.labels_full = adj_full, feat_full, raw.y
    d.idx_train, d.idx_val, d.idx_test = raw.idx_train, raw.idx_val, raw.idx_test
    it = raw.idx_train.numpy()
    d.adj_train = adj_full[np.ix_(it, it)]
    d.labels_train = raw.y[raw.idx_train]
    d.feat_train = feat_full[raw.idx_train]
    d.nclass = int(raw.num_classes)
    d.num_nodes = n
    return d


# ----------------------------------------------------------------------------- sparse side
class CsrBlock:
    """CSR matrix with fp32 values; ``matmul`` is differentiable w.r.t. the dense operand."""

    def __init__(self, rowptr, col, val, shape):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int64)
        self.val = np.ascontiguousarray(val, dtype=np.float32)
        self.shape = (int(shape[0]), int(shape[1]))
        self._t = None

    def t(self):
        if self._t is None:
            m = sp.csr_matrix((self.val, self.col, self.rowptr), shape=self.shape).T.tocsr()
            m.sort_indices()
            self._t = CsrBlock(m.indptr, m.indices, m.data, m.shape)
            self._t._t = self
        return self._t


class _CsrMatmul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, blk):
        ctx.blk = blk
        return torch.from_numpy(hostlib.spmm_csr(blk.rowptr, blk.col, blk.val, x.detach().numpy()))

    @staticmethod
    def backward(ctx, g):
        return _CsrMatmul.apply(g.contiguous(), ctx.blk.t()), None


def csr_matmul(blk, x):
    return _CsrMatmul.apply(x.contiguous(), blk)


def normalize_sparse(adj_csr):
    """utils.py:403-413 + :451-458.  float64 ``(D^-1/2 (A+I)) D^-1/2`` rounded once to fp32,
    returned as the CSR of its transpose (torch_sparse ``.t()``)."""
    a = sp.csr_matrix(adj_csr, dtype=np.float32)           # to_tensor(...).float() then to_scipy
    a = a + sp.eye(a.shape[0])                              # float64 from here on
    rowsum = np.array(a.sum(1))
    with np.errstate(divide="ignore"):
        r_inv = np.power(rowsum, -0.5).flatten()
    r_inv[np.isinf(r_inv)] = 0.0
    dm = sp.diags(r_inv)
    a = dm.dot(a).dot(dm)
    coo = a.tocoo().astype(np.float32)
    # coalesce (sorted by row, col), then transpose
    m = sp.csr_matrix((coo.data, (coo.row, coo.col)), shape=a.shape)
    m.sum_duplicates()
    mt = m.T.tocsr()
    mt.sort_indices()
    return CsrBlock(mt.indptr, mt.indices, mt.data, mt.shape)


def normalize_dense(adj):
    """utils.py:429-439."""
    mx = adj + torch.eye(adj.shape[0])
    rowsum = mx.sum(1)
    r_inv = rowsum.pow(-1 / 2).flatten()
    r_inv[torch.isinf(r_inv)] = 0.0
    r_mat = torch.diag(r_inv)
    return (r_mat @ mx) @ r_mat


def fanouts(dataset, nlayers):
    """loader.py:197-210."""
    if nlayers == 1:
        return [15]
    if nlayers == 2:
        return [15, 8] if dataset in ("reddit", "flickr") else [10, 5]
    return {3: [15, 10, 5], 4: [15, 10, 5, 5], 5: [15, 10, 5, 5, 5]}[nlayers]


class ClassSampler:
    """loader.py:187-224: per-class batch of <=256 train nodes + 2-hop sampled blocks."""

    def __init__(self, data, adj_norm, args):
        self.adj = adj_norm
        self.sizes = fanouts(args.dataset, args.nlayers)
        lt = data.labels_train.numpy()
        self.members = {}
        for c in range(data.nclass):
            if args.setting == "trans":
                self.members[c] = data.idx_train.numpy()[lt == c]
            else:
                self.members[c] = np.arange(len(lt))[lt == c]

    def sample(self, c, num=256):
        batch = np.random.permutation(self.members[c])[:num].astype(np.int64)
        n_id = batch
        blocks = []
        for k in self.sizes:
            rp, col, new_ids, e_id = hostlib.sample_adj(self.adj.rowptr, self.adj.col, n_id, k)
            blocks.append(CsrBlock(rp, col, self.adj.val[e_id], (n_id.size, new_ids.size)))
            n_id = new_ids
        return batch.size, n_id, blocks[::-1]


# ----------------------------------------------------------------------------- models
class _Affine(nn.Module):
    """MyLinear (layers.py:354-381) and the weight/bias of GraphConvolution (layers.py:17-34):
    weight (in,out) and bias both U(-1/sqrt(in), 1/sqrt(in))."""

    def __init__(self, fin, fout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(fin, fout))
        self.bias = nn.Parameter(torch.zeros(fout))
        self.reset_parameters()

    def reset_parameters(self):
        s = 1.0 / math.sqrt(self.weight.T.size(1))
        self.weight.data.uniform_(-s, s)
        self.bias.data.uniform_(-s, s)


class CondenseModel(nn.Module):
    """SGC (sgc.py:12-57) or GCN (gcn.py:8-23 + base.py:51-78) without BN / dropout."""

    def __init__(self, kind, nfeat, nhid, nclass, nlayers, ntrans):
        super().__init__()
        self.kind, self.nlayers = kind, nlayers
        if kind == "SGC":
            dims = [nfeat, nclass] if ntrans == 1 else [nfeat] + [nhid] * (ntrans - 1) + [nclass]
        elif kind == "GCN":
            dims = [nfeat, nclass] if nlayers == 1 else [nfeat] + [nhid] * (nlayers - 1) + [nclass]
        else:
            raise ValueError(kind)
        self.layers = nn.ModuleList([_Affine(a, b) for a, b in zip(dims[:-1], dims[1:])])

    def initialize(self):
        for layer in self.layers:
            layer.reset_parameters()

    @staticmethod
    def _prop(adj, x):
        return adj @ x if isinstance(adj, torch.Tensor) else csr_matmul(adj, x)

    def forward(self, x, adj):
        if self.kind == "SGC":
            for i, layer in enumerate(self.layers):
                x = x @ layer.weight + layer.bias
                if i != len(self.layers) - 1:
                    x = F.relu(x)
            for i in range(self.nlayers):
                x = self._prop(adj[i] if isinstance(adj, list) else adj, x)
        else:
            for i, layer in enumerate(self.layers):
                if isinstance(adj, list):
                    x = csr_matmul(adj[i], torch.mm(x, layer.weight)) + layer.bias
                else:
                    # layers.py:43-46: the dense path goes through a (1, N, out) view, i.e. a batched matmul
                    x = torch.mm(x.view(-1, x.shape[-1]), layer.weight)
                    x = adj @ x.view(-1, adj.shape[-1], x.shape[-1]) + layer.bias
                if i != self.nlayers - 1:
                    x = F.relu(x)
        return F.log_softmax(x.view(-1, x.shape[-1]), dim=1)


class PairwiseAdj(nn.Module):
    """PGE (parametrized_adj.py:7-86): MLP over all ordered node pairs, BN always in train mode."""

    def __init__(self, nfeat, nnodes, dataset, reduction_rate):
        super().__init__()
        nhid = 128
        if dataset in ("ogbn-arxiv", "arxiv", "flickr"):
            nhid = 256
        if dataset == "reddit":
            nhid = 128 if reduction_rate == 0.01 else 256
        self.layers = nn.ModuleList([nn.Linear(2 * nfeat, nhid), nn.Linear(nhid, nhid), nn.Linear(nhid, 1)])
        self.bns = nn.ModuleList([nn.BatchNorm1d(nhid), nn.BatchNorm1d(nhid)])
        for lin in self.layers:          # second draw: PGE.__init__ calls reset_parameters() again (:35)
            lin.reset_parameters()
        self.n = nnodes
        self.nchunks = 5 if (dataset == "reddit" and reduction_rate >= 0.01) else 1
        # np.meshgrid(arange, arange) then column_stack(X.ravel(), Y.ravel()): pair k=i*n+j -> (j, i)
        k = np.arange(nnodes * nnodes)
        self.e0, self.e1 = k % nnodes, k // nnodes

    def _mlp(self, h):
        for i, lin in enumerate(self.layers):
            h = lin(h)
            if i != len(self.layers) - 1:
                h = F.relu(self.bns[i](h))
        return h

    def forward(self, x):
        outs = []
        for idx in np.array_split(np.arange(self.e0.size), self.nchunks):
            outs.append(self._mlp(torch.cat([x[self.e0[idx]], x[self.e1[idx]]], dim=1)))
        adj = torch.cat(outs).reshape(self.n, self.n)
        adj = torch.sigmoid((adj + adj.T) / 2)
        return adj - torch.diag(torch.diag(adj, 0))

    @torch.no_grad()
    def inference(self, x):
        return self.forward(x)


# ----------------------------------------------------------------------------- matching loss
def _rowwise_cos_distance(gr, gs):
    if gr.dim() == 1:
        return 0
    gr, gs = gr.T, gs.T
    return torch.sum(1 - torch.sum(gr * gs, dim=-1) / (torch.norm(gr, dim=-1) * torch.norm(gs, dim=-1) + 0.000001))


def match_loss(gw_syn, gw_real, metric):
    """condensation/utils.py:12-106."""
    if metric == "ours":
        dis = torch.tensor(0.0)
        for gr, gs in zip(gw_real, gw_syn):
            dis = dis + _rowwise_cos_distance(gr, gs)
        return dis
    vr = torch.cat([g.reshape(-1) for g in gw_real])
    vs = torch.cat([g.reshape(-1) for g in gw_syn])
    if metric == "mse":
        return torch.sum((vs - vr) ** 2)
    if metric == "cos":
        return 1 - torch.sum(vr * vs, dim=-1) / (torch.norm(vr, dim=-1) * torch.norm(vs, dim=-1) + 0.000001)
    raise SystemExit("DC error: unknown distance function")


# ----------------------------------------------------------------------------- the reducer
def allocate_labels(labels_train, rate):
    """gcond_base.py:79-115 -> (labels_syn, num_class_dict in allocation order)."""
    counts = Counter(labels_train.tolist())
    n = len(labels_train)
    ordered = sorted(counts.items(), key=lambda kv: kv[1])
    alloc, used, labels = {}, 0, []
    for i, (c, num) in enumerate(ordered):
        if i == len(ordered) - 1:
            alloc[c] = max(int(n * rate) - used, 1)
        else:
            alloc[c] = max(int(num * rate), 1)
            used += alloc[c]
        labels += [c] * alloc[c]
    return np.array(labels), alloc


def random_init_ids(data, alloc, setting):
    """random.py:9-17 with coreset_base.py:14-21 index conventions."""
    lt = data.labels_train.numpy()
    base = np.arange(len(lt)) if setting == "ind" else data.idx_train.numpy()
    picks = []
    for c, cnt in alloc.items():
        picks.append(np.random.permutation(base[lt == c])[:cnt])
    return np.hstack(picks)


class GCondOracle:
    """Mirrors ``GCond(setting, data, args)`` / ``GCondX`` construction + ``reduce``.

    ``observer`` (optional) receives events: ('norm', CsrBlock), ('sample', step, c, bs, n_id, blocks),
    ('loss', step, float), ('grads', step, feat_grad, pge_grads), ('model_init', epoch, flat_params).
    """

    def __init__(self, data, args, observer=None):
        self.data, self.args = data, args
        self.x_variant = args.method in ("gcondx", "doscondx")
        # DosCond / DosCondX (condensation/doscond.py:45-58, doscondx.py:44-53): one matching step per outer step,
        # every optimiser steps every time, the condense model is never trained
        self.one_step = args.method in ("doscond", "doscondx")
        self.obs = observer or (lambda *a: None)
        self.labels_syn_np, self.alloc = allocate_labels(data.labels_train, args.reduction_rate)
        self.n_syn = n = self.labels_syn_np.shape[0]
        self.d = d = data.feat_train.shape[1]
        self.feat_syn = nn.Parameter(torch.empty(n, d))
        self.pge = PairwiseAdj(d, n, args.dataset, args.reduction_rate)
        self.opt_feat = torch.optim.Adam([self.feat_syn], lr=args.lr_feat)
        self.opt_pge = torch.optim.Adam(self.pge.parameters(), lr=args.lr_adj)
        self.adj_syn = None

    # gcond_base.py:156-241
    def _match_all_classes(self, model, sampler, features, labels, labels_syn, step):
        args = self.args
        loss = torch.tensor(0.0)
        params = list(model.parameters())
        for c in range(self.data.nclass):
            bs, n_id, blocks = sampler.sample(c)
            self.obs("sample", step, c, bs, n_id, blocks)
            out_real = model(features[torch.from_numpy(n_id)], blocks)
            loss_real = F.nll_loss(out_real, labels[torch.from_numpy(n_id[:bs])])
            gw_real = [g.detach().clone() for g in torch.autograd.grad(loss_real, params)]
            out_syn = model(self.feat_syn, self.adj_syn)
            sel = labels_syn == c
            loss_syn = F.nll_loss(out_syn[sel], labels_syn[sel])
            gw_syn = torch.autograd.grad(loss_syn, params, create_graph=True)
            loss = loss + (self.alloc[c] / self.n_syn) * match_loss(gw_syn, gw_real, args.dis_metric)
        return loss

    def reduce(self, epochs=None, max_outer_steps=None):
        """max_outer_steps: stop after that many outer steps (bounded CPU-baseline samples in bench.py)."""
        import time
        args, data = self.args, self.data
        labels_syn = torch.from_numpy(self.labels_syn_np).long()
        if args.setting == "trans":
            features, adj_sp, labels = data.feat_full.float(), data.adj_full, data.labels_full.long()
        else:
            features, adj_sp, labels = data.feat_train.float(), data.adj_train, data.labels_train.long()
        ids = random_init_ids(data, self.alloc, args.setting)
        self.init_ids = ids
        src = data.feat_full if args.setting == "trans" else data.feat_train
        if getattr(args, "agg", False):
            # model_free_coreset_base.py:18-27: features aggregated over two hops, (A_hat A_hat) X with the sparse-sparse
            # product first (trans only: the reference's 'ind' branch multiplies a train-sized operator with feat_full)
            a = normalize_sparse(data.adj_full)
            m = sp.csr_matrix((a.val, a.col, a.rowptr), shape=a.shape)
            p2 = (m @ m).tocsr().astype(np.float32)
            p2.sort_indices()
            src = torch.from_numpy(hostlib.spmm_csr(p2.indptr.astype(np.int64), p2.indices.astype(np.int64), p2.data,
                                                    data.feat_full.float().numpy()))
        self.feat_syn.data.copy_(src[torch.from_numpy(ids)].float())
        if self.x_variant:
            self.adj_syn = torch.eye(self.n_syn)
        adj = normalize_sparse(adj_sp)
        self.obs("norm", adj)
        sampler = ClassSampler(data, adj, args)
        model = CondenseModel(args.condense_model, self.d, args.hidden, data.nclass, args.nlayers, args.ntrans)
        losses = []
        step = 0
        self.step_done_at = []
        self.loop_started_at = time.perf_counter()
        for it in range(args.epochs if epochs is None else epochs):
            model.initialize()
            self.obs("model_init", it, torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy())
            opt_model = torch.optim.Adam(model.parameters(), lr=args.lr)
            for ol in range(args.outer_loop):
                if not self.x_variant:
                    self.adj_syn = normalize_dense(self.pge(self.feat_syn))
                loss = self._match_all_classes(model, sampler, features, labels, labels_syn, step)
                losses.append(float(loss.item()))
                self.obs("loss", step, losses[-1])
                self.opt_feat.zero_grad()
                self.opt_pge.zero_grad()
                loss.backward()
                self.obs("grads", step, self.feat_syn.grad,
                         [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.pge.parameters()])
                if self.one_step:
                    if not self.x_variant:
                        self.opt_pge.step()
                    self.opt_feat.step()
                    step += 1
                    self.step_done_at.append(time.perf_counter())
                    if max_outer_steps is not None and step >= max_outer_steps:
                        self.losses = losses
                        return losses
                    continue
                pge_turn = (ol % 5 < 1) if self.x_variant else (it % 50 < 10)
                (self.opt_pge if pge_turn else self.opt_feat).step()
                step += 1
                feat_inner = self.feat_syn.detach()
                if self.x_variant:
                    adj_inner = self.adj_syn
                else:
                    self.adj_inner_raw = self.pge.inference(feat_inner)
                    adj_inner = normalize_dense(self.adj_inner_raw)
                for _ in range(args.inner_loop):
                    opt_model.zero_grad()
                    F.nll_loss(model(feat_inner, adj_inner), labels_syn).backward()
                    opt_model.step()
                self.step_done_at.append(time.perf_counter())
                if max_outer_steps is not None and step >= max_outer_steps:
                    self.losses = losses
                    return losses
        self.losses = losses
        return losses

    def result(self):
        """What the reference writes to data.* at a checkpoint (gcond.py:76-78 / gcondx.py:74-76)."""
        if self.one_step and not self.x_variant:
            self.adj_inner_raw = self.pge.inference(self.feat_syn.detach())        # doscond.py:61
        adj = torch.eye(self.n_syn) if self.x_variant else self.adj_inner_raw.detach()
        return adj, self.feat_syn.detach(), torch.from_numpy(self.labels_syn_np).long()
