"""PGE: pairwise adjacency generator (graphslim/models/parametrized_adj.py:7-86), forward and backward.

Parameters live as plain device tensors in nn.Linear / BatchNorm1d layout and order so that
``parameters()`` lines up with the reference's ``pge.parameters()``:
    layers.0.weight (h,2d), layers.0.bias, layers.1.weight (h,h), layers.1.bias, layers.2.weight (1,h),
    layers.2.bias, bns.0.weight, bns.0.bias, bns.1.weight, bns.1.bias.

Initial values are drawn with torch's CPU generator in the reference's order (each Linear twice: once in its
constructor, once more by PGE.reset_parameters, parametrized_adj.py:35,79-86) and copied to HBM, so a run
is seed-comparable with the reference's gpu_id=-1 path.
"""
import numpy as np
import torch


def pge_hidden(dataset, reduction_rate):
    """parametrized_adj.py:11-17."""
    nhid = 128
    if dataset in ("ogbn-arxiv", "arxiv", "flickr"):
        nhid = 256
    if dataset in ("reddit",):
        nhid = 128 if reduction_rate == 0.01 else 256
    return nhid


class PGE:
    def __init__(self, K, nfeat, nnodes, args):
        self.K = K
        self.n, self.d = int(nnodes), int(nfeat)
        self.h = h = pge_hidden(args.dataset, args.reduction_rate)
        self.nchunks = 5 if (args.dataset == "reddit" and args.reduction_rate >= 0.01) else 1
        dims = [(2 * nfeat, h), (h, h), (h, 1)]
        lins = [torch.nn.Linear(i, o) for i, o in dims]          # first draw (constructor)
        for lin in lins:                                          # second draw (PGE.reset_parameters)
            lin.reset_parameters()
        dev = K.device
        self.W = [lin.weight.detach().clone().to(dev) for lin in lins]
        self.b = [lin.bias.detach().clone().to(dev) for lin in lins]
        self.gamma = [torch.ones(h, device=dev), torch.ones(h, device=dev)]
        self.beta = [torch.zeros(h, device=dev), torch.zeros(h, device=dev)]
        # np.array_split boundaries over the n*n pair rows (parametrized_adj.py:43)
        total = self.n * self.n
        sizes = [total // self.nchunks + (1 if i < total % self.nchunks else 0) for i in range(self.nchunks)]
        self.chunk_off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64, device=dev)
        self.eps = 1e-5

    # ---------------------------------------------------------------------------------- row sharding
    def enable_row_sharding(self, group=None):
        """Deal the N'^2 pair rows (i, j) to the ranks of `group` by contiguous slices of i ("PGE's N'^2
        pair rows can additionally be row-sharded with an allreduce of BN partial sums and an allgather of adjacency
        rows").  Parameters stay replicated; every rank ends a forward with the full adjacency and a backward with the
        full gradients, so the optimiser steps remain bit-identical across ranks.  The per-chunk BatchNorm of the
        reddit >= 0.01 configuration is left replicated."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world == 1 or self.nchunks != 1 or self.n < world:
            self.shard = None
            return False
        n, dev = self.n, self.K.device
        bounds = [n * r // world for r in range(world + 1)]
        rows = [(bounds[r + 1] - bounds[r]) * n for r in range(world)]
        self.shard = dict(group=group, world=world, rank=rank, i0=bounds[rank], i1=bounds[rank + 1], bounds=bounds,
                          rows=rows, pad=max(rows),
                          off_rows=torch.tensor([0, rows[rank]], dtype=torch.int64, device=dev),
                          counts=torch.tensor(rows, dtype=torch.int64, device=dev))
        return True

    def _forward_sharded(self, x, keep):
        import torch.distributed as dist
        K, n, d, h, sh = self.K, self.n, self.d, self.h, self.shard
        W1 = self.W[0]
        Pa = K.gemm(x, W1[:, :d], tb=True)
        Pb = K.gemm(x, W1[:, d:], tb=True)
        mean1, rstd1, cm1 = K.pge_l1_stats_closed(Pa, Pb, self.eps)            # 2n rows: replicated
        H1 = K.pge_l1_expand_rows(Pa, Pb[sh["i0"]:sh["i1"]], sh["off_rows"], mean1, rstd1, self.gamma[0], self.beta[0])
        with K.timed("pge_l2_fwd"):
            Y2 = K.gemm(H1, self.W[1], tb=True)
        part = torch.cat([K.col_stats_partial(Y2, sh["off_rows"]), Y2[0].double()])
        parts = torch.empty(sh["world"] * 3 * h, dtype=torch.float64, device=K.device)
        dist.all_gather_into_tensor(parts, part, group=sh["group"])            # BN2 partial sums: 3h doubles per rank
        mean2, rstd2 = K.col_stats_combine(parts.view(sh["world"], 3 * h), sh["counts"], self.eps)
        E_loc = K.pge_l3(Y2, sh["off_rows"], mean2, rstd2, self.gamma[1], self.beta[1], self.W[2].view(-1), self.b[2])
        send = E_loc
        if E_loc.numel() != sh["pad"]:
            send = torch.zeros(sh["pad"], dtype=torch.float32, device=K.device)
            send[:E_loc.numel()] = E_loc
        full = torch.empty(sh["world"] * sh["pad"], dtype=torch.float32, device=K.device)
        dist.all_gather_into_tensor(full, send, group=sh["group"])             # adjacency rows of every rank
        if all(r == sh["pad"] for r in sh["rows"]):
            E = full
        else:
            E = torch.cat([full[r * sh["pad"]: r * sh["pad"] + sh["rows"][r]] for r in range(sh["world"])])
        A = K.pge_symm_sigmoid(E, n)
        if keep:
            self._saved = (x, Pa, Pb, mean1, rstd1, cm1, H1, Y2, mean2, rstd2, A)
        return A

    def _backward_sharded(self, dA):
        import torch.distributed as dist
        K, n, d, h, sh = self.K, self.n, self.d, self.h, self.shard
        x, Pa, Pb, mean1, rstd1, cm1, H1, Y2, mean2, rstd2, A = self._saved
        W1, W2, w3 = self.W[0], self.W[1], self.W[2].view(-1)
        grp = sh["group"]
        dE = K.pge_symm_sigmoid_bwd(dA, A)
        dE_loc = dE[sh["i0"] * n: sh["i1"] * n]
        s1, s2, dw3, db3 = K.pge_l3_bwd_stats(Y2, dE_loc, sh["off_rows"], mean2, rstd2, self.gamma[1], self.beta[1], w3)
        flat = torch.cat([s1.view(-1), s2.view(-1), dw3.view(-1), db3.view(-1)])
        dist.all_reduce(flat, group=grp)                                       # BN2 backward sums + layer-3 grads
        s1, s2, dw3, db3 = flat[:h].view(1, h), flat[h:2 * h].view(1, h), flat[2 * h:3 * h], flat[3 * h:3 * h + 1]
        dgamma2, dbeta2 = s2.sum(0), s1.sum(0)
        # the apply kernel divides the sums by the row count of the rows it is given: pre-scale to the global count
        scale = float(sh["rows"][sh["rank"]]) / float(n * n)
        dY2 = K.pge_bn2_bwd_apply(Y2, dE_loc, sh["off_rows"], mean2, rstd2, self.gamma[1], self.beta[1], w3,
                                  s1 * scale, s2 * scale)
        with K.timed("pge_l2_bwd_dw"):
            dW2 = K.gemm(dY2, H1, ta=True)
        with K.timed("pge_l2_bwd_dx"):
            dH1 = K.gemm(dY2, W2)
        work = K.pge_bn1_bwd_pass_rows(dH1, Pa, Pb, sh["i0"], sh["i1"] - sh["i0"], mean1, rstd1, self.gamma[0],
                                       self.beta[0])
        dist.all_reduce(work[:2 * h], group=grp)                               # t1, t2 (float64)
        red = torch.cat([work[2 * h:].view(torch.float32), dW2.view(-1)])
        dist.all_reduce(red, group=grp)                                        # Ga, Gb and dW2 (float32)
        nfl = 2 * n * h
        work[2 * h:].view(torch.float32).copy_(red[:nfl])
        dW2 = red[nfl:].view(h, h)
        dPa, dPb, dgamma1, dbeta1 = K.pge_bn1_bwd_final(Pa, Pb, rstd1, self.gamma[0], cm1, work)
        dW1 = K.empty(h, 2 * d)
        K.gemm(dPa, x, ta=True, out=dW1[:, :d])
        K.gemm(dPb, x, ta=True, out=dW1[:, d:])
        dX = K.gemm(dPa, W1[:, :d])
        K.gemm(dPb, W1[:, d:], out=dX, beta=1.0)
        zeros_h = K.zeros(h)
        grads = [dW1, zeros_h, dW2, zeros_h.clone(), dw3.reshape(1, h), db3.reshape(1), dgamma1, dbeta1, dgamma2, dbeta2]
        self._saved = None
        return grads, dX

    # ---------------------------------------------------------------------------------- fused layer-2 pipeline
    def _slice(self):
        """(shard dict or None, first i, number of i) of this rank's pair rows."""
        sh = getattr(self, "shard", None)
        if sh is None:
            return None, 0, self.n
        return sh, sh["i0"], sh["i1"] - sh["i0"]

    def _forward_fused(self, x, keep):
        """csrc/pge_fused.cu: H1 is generated inside the layer-2 product's A-producer and the BatchNorm-2 column sums
        come out of its epilogue, so one N'^2 x h array (Y2) is written and read once instead of 5 passes.  With row
        sharding the same kernels run on the rank's slice of i; the column sums are all-reduced (2h doubles)."""
        K, n, d, h = self.K, self.n, self.d, self.h
        sh, i0, n_i = self._slice()
        W1 = self.W[0]
        Pa = K.gemm(x, W1[:, :d], tb=True)
        Pb = K.gemm(x, W1[:, d:], tb=True)
        mean1, rstd1, cm1 = K.pge_l1_stats_closed(Pa, Pb, self.eps)
        with K.timed("pge_l2_fwd"):
            Y2, stats = K.pge_fused_l2_fwd(Pa, Pb, i0, n_i, mean1, rstd1, self.gamma[0], self.beta[0], self.W[1])
        if sh is not None:
            import torch.distributed as dist
            dist.all_reduce(stats, group=sh["group"])
        mean2, rstd2 = K.pge_stats_finalize(stats, float(n) * float(n), self.eps)
        off = self.chunk_off if sh is None else sh["off_rows"]
        E = K.pge_l3(Y2, off, mean2, rstd2, self.gamma[1], self.beta[1], self.W[2].view(-1), self.b[2])
        if sh is not None:
            E = self._gather_rows(E, sh)
        A = K.pge_symm_sigmoid(E, n)
        if keep:
            self._saved = (x, Pa, Pb, mean1, rstd1, cm1, None, Y2, mean2, rstd2, A)
        return A

    def _gather_rows(self, E_loc, sh):
        """All-gather of the adjacency rows of every rank (slices may be uneven: padded to the largest)."""
        import torch.distributed as dist
        K = self.K
        send = E_loc
        if E_loc.numel() != sh["pad"]:
            send = torch.zeros(sh["pad"], dtype=torch.float32, device=K.device)
            send[:E_loc.numel()] = E_loc
        full = torch.empty(sh["world"] * sh["pad"], dtype=torch.float32, device=K.device)
        dist.all_gather_into_tensor(full, send, group=sh["group"])
        if all(r == sh["pad"] for r in sh["rows"]):
            return full
        return torch.cat([full[r * sh["pad"]: r * sh["pad"] + sh["rows"][r]] for r in range(sh["world"])])

    def _backward_fused(self, dA, need_params=True, need_dx=True):
        """Backward of `_forward_fused`: dY2 is recomputed from Y2 inside the producers of both layer-2 products, dH1 is
        masked and reduced in the epilogue of the first, dW2 accumulates in TMEM in the second."""
        K, n, d, h = self.K, self.n, self.d, self.h
        sh, i0, n_i = self._slice()
        x, Pa, Pb, mean1, rstd1, cm1, _, Y2, mean2, rstd2, A = self._saved
        W1, W2, w3 = self.W[0], self.W[1], self.W[2].view(-1)
        bn1 = (mean1, rstd1, self.gamma[0], self.beta[0])
        bn2 = (mean2, rstd2, self.gamma[1], self.beta[1])
        count = float(n) * float(n)
        dE = K.pge_symm_sigmoid_bwd(dA, A)
        off = self.chunk_off
        if sh is not None:
            dE = dE[i0 * n: (i0 + n_i) * n]
            off = sh["off_rows"]
        s1, s2, dw3, db3 = K.pge_l3_bwd_stats(Y2, dE, off, mean2, rstd2, self.gamma[1], self.beta[1], w3)
        if sh is not None:
            import torch.distributed as dist
            flat = torch.cat([s1.view(-1), s2.view(-1), dw3.view(-1), db3.view(-1)])
            dist.all_reduce(flat, group=sh["group"])                          # BN2 backward sums + layer-3 grads
            s1, s2, dw3, db3 = flat[:h].view(1, h), flat[h:2 * h].view(1, h), flat[2 * h:3 * h], flat[3 * h:3 * h + 1]
        dgamma2, dbeta2 = s2.sum(0), s1.sum(0)
        work = K.pge_bn1_work(n, h)
        with K.timed("pge_l2_bwd_dx"):
            K.pge_fused_l2_bwd_dx(Pa, Pb, i0, n_i, bn1, W2, Y2, dE, bn2, w3, s1, s2, count, work=work)
        dW2 = None
        if need_params:
            with K.timed("pge_l2_bwd_dw"):
                dW2 = K.pge_fused_l2_bwd_dw(Pa, Pb, i0, n_i, bn1, Y2, dE, bn2, w3, s1, s2, count)
        if sh is not None:
            # one collective: Ga, Gb (and dW2), float32.  The BN1 sums t1, t2 are linear in Ga / Gb, so they are formed
            # from the reduced tiles afterwards (replicated, tiny) instead of travelling as a second all-reduce
            nfl = 2 * n * h
            flt = work[2 * h:].view(torch.float32)
            red = torch.cat([flt, dW2.view(-1)]) if need_params else flt
            dist.all_reduce(red, group=sh["group"])
            if need_params:
                flt.copy_(red[:nfl])
                dW2 = red[nfl:].view(h, h)
        K.pge_bn1_tsum(Pa, Pb, cm1, rstd1, work)
        dPa, dPb, dgamma1, dbeta1 = K.pge_bn1_bwd_final(Pa, Pb, rstd1, self.gamma[0], cm1, work)
        grads = dX = None
        if need_params:
            dW1 = K.empty(h, 2 * d)
            K.gemm(dPa, x, ta=True, out=dW1[:, :d])
            K.gemm(dPb, x, ta=True, out=dW1[:, d:])
            zeros_h = K.zeros(h)
            grads = [dW1, zeros_h, dW2, zeros_h.clone(), dw3.reshape(1, h), db3.reshape(1), dgamma1, dbeta1, dgamma2,
                     dbeta2]
        if need_dx:
            dX = K.gemm(dPa, W1[:, :d])
            K.gemm(dPb, W1[:, d:], out=dX, beta=1.0)
        self._saved = None
        return grads, dX

    def parameters(self):
        return [self.W[0], self.b[0], self.W[1], self.b[1], self.W[2], self.b[2],
                self.gamma[0], self.beta[0], self.gamma[1], self.beta[1]]

    # ---------------------------------------------------------------------------------- forward
    def forward(self, x, keep=True):
        """adj (n,n) = zero-diag(sigmoid((E+E^T)/2)),  E[i,j] = MLP([x_j, x_i]).  BN always uses batch stats."""
        if self.K.pge_fused_supported(self.h, self.nchunks):
            return self._forward_fused(x, keep)
        if getattr(self, "shard", None) is not None:
            return self._forward_sharded(x, keep)
        K, n, d, h = self.K, self.n, self.d, self.h
        W1 = self.W[0]
        Pa = K.gemm(x, W1[:, :d], tb=True)                       # layer 1, first half: indexed by j
        Pb = K.gemm(x, W1[:, d:], tb=True)                       # second half: indexed by i (bias cancels in BN)
        cm1 = None
        if self.nchunks == 1:                                     # product-set statistics factorise: 2n rows, not n^2
            mean1, rstd1, cm1 = K.pge_l1_stats_closed(Pa, Pb, self.eps)
        else:
            mean1, rstd1 = K.pge_l1_stats(Pa, Pb, self.chunk_off, self.eps)
        H1 = K.pge_l1_expand(Pa, Pb, self.chunk_off, mean1, rstd1, self.gamma[0], self.beta[0])
        with K.timed("pge_l2_fwd"):
            Y2 = K.gemm(H1, self.W[1], tb=True)                  # the N'^2 x h x h product (bias cancels in BN)
        mean2, rstd2 = K.col_stats_chunked(Y2, self.chunk_off, self.eps)
        E = K.pge_l3(Y2, self.chunk_off, mean2, rstd2, self.gamma[1], self.beta[1], self.W[2].view(-1), self.b[2])
        A = K.pge_symm_sigmoid(E, n)
        if keep:
            self._saved = (x, Pa, Pb, mean1, rstd1, cm1, H1, Y2, mean2, rstd2, A)
        return A

    def inference(self, x, keep=False):
        """parametrized_adj.py:73-77: same forward without autograd (BN still in train mode).  `keep=True` retains the
        activations: nothing changes PGE's parameters or feat_syn between this call and the next outer step's
        `pge(feat_syn)` (gcond.py:63 -> :48; the inner loop only trains the condense model), so that forward would
        recompute exactly these tensors and the caller may back-propagate through this pass instead."""
        return self.forward(x, keep=keep)

    # ---------------------------------------------------------------------------------- backward
    def backward(self, dA, need_params=True, need_dx=True):
        """Returns (grads in parameters() order, dX).  `need_params=False` / `need_dx=False` let the caller skip the half
        of the backward whose result it will discard: the reference back-propagates into both the PGE parameters and
        feat_syn on every outer step but steps only one optimiser (gcond.py:54-61) and zeroes the other gradient
        unread, so the N'^2-deep dW2 product (and the layer-1 weight products) are dead work on feature turns.  The
        fused path honours the flags (returns None for the skipped half); the other paths compute everything."""
        if self._saved[6] is None:                               # saved by the fused forward: H1 was never materialised
            return self._backward_fused(dA, need_params, need_dx)
        if getattr(self, "shard", None) is not None:
            return self._backward_sharded(dA)
        K, n, d, h = self.K, self.n, self.d, self.h
        x, Pa, Pb, mean1, rstd1, cm1, H1, Y2, mean2, rstd2, A = self._saved
        W1, W2, w3 = self.W[0], self.W[1], self.W[2].view(-1)
        dE = K.pge_symm_sigmoid_bwd(dA, A)
        s1, s2, dw3, db3 = K.pge_l3_bwd_stats(Y2, dE, self.chunk_off, mean2, rstd2, self.gamma[1], self.beta[1], w3)
        dgamma2, dbeta2 = s2.sum(0), s1.sum(0)
        dY2 = K.pge_bn2_bwd_apply(Y2, dE, self.chunk_off, mean2, rstd2, self.gamma[1], self.beta[1], w3, s1, s2)
        with K.timed("pge_l2_bwd_dw"):
            dW2 = K.gemm(dY2, H1, ta=True)                        # (h_out, h_in), K = N'^2
        with K.timed("pge_l2_bwd_dx"):
            dH1 = K.gemm(dY2, W2)                                 # N'^2 x h
        if cm1 is not None:                                       # one pass over dH1 (linear reductions only)
            dPa, dPb, dgamma1, dbeta1 = K.pge_bn1_bwd_closed(dH1, Pa, Pb, mean1, rstd1, self.gamma[0], self.beta[0],
                                                             cm1)
        else:
            t1, t2 = K.pge_bn1_bwd_stats(dH1, Pa, Pb, self.chunk_off, mean1, rstd1, self.gamma[0], self.beta[0])
            dgamma1, dbeta1 = t2.sum(0), t1.sum(0)
            dPa, dPb = K.pge_bn1_bwd_reduce(dH1, Pa, Pb, self.chunk_off, mean1, rstd1, self.gamma[0], self.beta[0],
                                            t1, t2)
        dW1 = K.empty(h, 2 * d)
        K.gemm(dPa, x, ta=True, out=dW1[:, :d])
        K.gemm(dPb, x, ta=True, out=dW1[:, d:])
        dX = K.gemm(dPa, W1[:, :d])
        K.gemm(dPb, W1[:, d:], out=dX, beta=1.0)
        zeros_h = K.zeros(h)                                      # biases ahead of a BatchNorm get exactly zero gradient
        grads = [dW1, zeros_h, dW2, zeros_h.clone(), dw3.view(1, h), db3, dgamma1, dbeta1, dgamma2, dbeta2]
        self._saved = None
        return grads, dX
