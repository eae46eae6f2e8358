"""Closed-form gradient matching for the condense models (SGC / GCN), all classes at once.

The reference (graphslim/condensation/gcond_base.py:156-241) loops over classes; for each class it
runs the model on the sampled real block, takes d loss / d theta, runs the *same* synthetic forward
again, takes d loss_c / d theta with create_graph=True, and lets autograd differentiate the matching
distance a second time.  Here the whole thing is a fixed sequence of dense products:

  * one synthetic forward shared by all classes;
  * per-class first-order gradients kept side by side in "class-column" matrices
    (rows x n_class*width; column block c belongs to class c);
  * the second-order part as a forward tangent pass along G_c = coeff_c * dD/dg_c followed by one
    reverse pass (d/dX and d/dA_hat of  sum_c <G_c, grad_theta loss_c>), see DESIGN.md section 4.

`K` is the kernel namespace (graphslim_b200.ops.CudaOps).  Nothing here touches autograd.
"""
import torch


def _ones_col(K, n):
    return torch.ones(n, 1, dtype=torch.float32, device=K.device)


class ClassLayout:
    """Row -> class bookkeeping of the synthetic nodes (labels_syn is a run of classes)."""

    def __init__(self, K, labels_syn, n_class, owned=None):
        """owned: class ids this rank matches (class sharding); None = every class.  Class-column matrices carry one
        block per owned class; rows of other classes get block index -1 and contribute nothing."""
        dev = K.device
        lab = torch.as_tensor(labels_syn, dtype=torch.int64)
        self.n = lab.numel()
        self.n_class = int(n_class)
        self.owned = list(range(self.n_class)) if owned is None else sorted(int(c) for c in owned)
        self.nblk = len(self.owned)
        local = torch.full((self.n_class,), -1, dtype=torch.int64)
        local[torch.tensor(self.owned, dtype=torch.int64)] = torch.arange(self.nblk)
        counts = torch.bincount(lab, minlength=self.n_class).to(torch.float32)
        self.labels = lab.to(torch.int32).to(dev)
        self.blk = local[lab].to(torch.int32).to(dev)           # row -> class-column block (or -1)
        self.inv_nc_row = (1.0 / counts[lab]).to(dev)          # 1/n_c per synthetic row (nll mean over the class)
        self.coeff = (counts / float(self.n))[torch.tensor(self.owned, dtype=torch.int64)].to(dev)   # n_c / N'
        self.inv_n_row = torch.full((self.n,), 1.0 / self.n, dtype=torch.float32, device=dev)
        self.mask = None if owned is None else [1 if c in set(self.owned) else 0 for c in range(self.n_class)]


class RealBatch:
    """Device view of one outer step's sampled blocks (see sampler.py)."""
    pass


class _ModelBase:
    def __init__(self, K, d, hidden, n_class, nlayers, layout, identity_adj=False):
        self.K, self.d, self.h, self.C, self.k = K, d, hidden, n_class, nlayers
        self.lay = layout
        self.identity_adj = identity_adj
        self.W = None

    # propagate with the dense synthetic adjacency (or identity for GCondX)
    def _prop(self, A, M, transpose=False, fresh=False, bias=None, relu=False):
        """A M (or A^T M), optionally followed by + bias and ReLU inside the product's store."""
        if self.identity_adj:
            epi = bias is not None or relu
            out = M.clone() if (fresh or epi) else M  # `fresh`: the caller will modify the result in place
            return self.K.bias_act(out, bias, relu=relu) if epi else out
        return self.K.gemm(A, M, ta=transpose, bias=bias, relu=relu)

    def set_weights(self, W):
        self.W = W

    def _acc_dA(self, dA, L, R):
        """dA += L @ R^T (skipped when the adjacency is the fixed identity)."""
        if self.identity_adj:
            return dA
        if dA is None:
            return self.K.gemm(L, R, tb=True)
        return self.K.gemm(L, R, tb=True, out=dA, beta=1.0)


# ======================================================================================== SGC, ntrans = 1
class SGC1(_ModelBase):
    """Z = A^k (X W + 1 b^T)   (graphslim/models/sgc.py:37-57 with ntrans == 1)."""
    param_shapes = property(lambda s: [(s.d, s.C), (s.C,)])
    widths = property(lambda s: [s.C, s.C])
    is_bias = [False, True]

    # ---- real side: H^r = A1 (A2 X[n_id]) does not depend on W, so propagate at feature width once
    def real_grads(self, rb, X_full, ones_full, X_padded=None, out=None):
        """`out`: optional list of fixed result buffers (one per parameter, class-column layout), cleared and written in
        place of fresh allocations (the matching graph's input buffers)."""
        K = self.K
        W, b = self.W
        o = out if out is not None else [None] * 2
        d = X_full.shape[1]
        Xp = X_full if X_padded is None else X_padded      # rows padded with zeros to a float4 multiple
        T = K.spmm(rb.blocks_fwd[0].with_global_cols(), Xp)
        t = K.spmm(rb.blocks_fwd[0].with_global_cols(), ones_full)
        for blk in rb.blocks_fwd[1:]:
            T = K.spmm(blk.csr, T)
            t = K.spmm(blk.csr, t)
        T = T[:, :d]
        Z = K.gemm(T, W)
        K.gemm(t, b.view(1, -1), out=Z, beta=1.0)
        _, R = K.softmax_residual(Z, rb.labels, rb.inv_b)
        gW = K.gemm_grouped_tn(T, R, rb.seg[0], rb.out_block, self.lay.nblk, aligned=rb.aligned, out=o[0])
        gb = K.gemm_grouped_tn(t, R, rb.seg[0], rb.out_block, self.lay.nblk, aligned=rb.aligned, out=o[1])
        return [gW, gb]

    def syn_forward(self, X, A):
        K = self.K
        W, b = self.W
        self.X, self.A = X, A
        self.T = [X]
        self.t = [_ones_col(K, X.shape[0])]
        for _ in range(self.k):
            self.T.append(self._prop(A, self.T[-1]))
            self.t.append(self._prop(A, self.t[-1]))
        H, u = self.T[-1], self.t[-1]
        Z = K.gemm(H, W)
        K.gemm(u, b.view(1, -1), out=Z, beta=1.0)
        self.Z = Z
        return Z

    def syn_grads(self):
        K, lay = self.K, self.lay
        self.S, self.R = K.softmax_residual(self.Z, lay.labels, lay.inv_nc_row)
        self.Rexp = K.expand_class_blocks(self.R, lay.blk, lay.nblk)
        gW = K.gemm(self.T[-1], self.Rexp, ta=True)
        gb = K.gemm(self.t[-1], self.Rexp, ta=True)
        return [gW, gb]

    def syn_backward(self, G, need_dA=True):
        K, lay = self.K, self.lay
        W, b = self.W
        GW, Gb = G
        H, u = self.T[-1], self.t[-1]
        Zt = K.gemm(H, GW)
        K.gemm(u, Gb, out=Zt, beta=1.0)
        q = K.pick_class_blocks(Zt, lay.blk, lay.nblk)
        dZ = K.softmax_jvp(self.S, q, lay.inv_nc_row)
        dT = K.gemm(self.Rexp, GW, tb=True)
        K.gemm(dZ, W, tb=True, out=dT, beta=1.0)
        dt = K.gemm(self.Rexp, Gb, tb=True)
        K.gemm(dZ, b.view(1, -1), tb=True, out=dt, beta=1.0)
        dA = None
        for l in range(self.k, 0, -1):
            if need_dA:
                dA = self._acc_dA(dA, dT, self.T[l - 1])
                dA = self._acc_dA(dA, dt, self.t[l - 1])
            dT = self._prop(self.A, dT, transpose=True)
            dt = self._prop(self.A, dt, transpose=True)
        return dT, dA

    # ---- inner loop: plain training gradients on (X, A) with labels_syn, nll mean over all nodes
    def train_grads(self, X, A):
        K, lay = self.K, self.lay
        Z = self.syn_forward(X, A)
        _, R = K.softmax_residual(Z, lay.labels, lay.inv_n_row)
        return [K.gemm(self.T[-1], R, ta=True), K.gemm(self.t[-1], R, ta=True).view(-1)]


# ======================================================================================== SGC, ntrans = 2
class SGC2(_ModelBase):
    """Z = A^k (relu(X W1 + b1) W2 + b2)   (sgc.py:37-57 with ntrans == 2)."""
    param_shapes = property(lambda s: [(s.d, s.h), (s.h,), (s.h, s.C), (s.C,)])
    widths = property(lambda s: [s.h, s.h, s.C, s.C])
    is_bias = [False, True, False, True]

    def real_grads(self, rb, X_full, ones_full, X_padded=None, out=None):
        K = self.K
        W1, b1, W2, b2 = self.W
        o = out if out is not None else [None] * 4
        Xg = K.gather_rows(X_full, rb.nid)
        H1 = K.gemm(Xg, W1, bias=b1, relu=True)
        U = K.gemm(H1, W2, bias=b2)
        T = U
        for blk in rb.blocks_fwd:
            T = K.spmm(blk.csr, T)
        _, R = K.softmax_residual(T, rb.labels, rb.inv_b)
        dU = R
        for blk in reversed(rb.blocks_fwd):
            dU = K.spmm(blk.csr_t, dU)
        seg, ids, nb = rb.seg[-1], rb.out_block, self.lay.nblk
        gW2 = K.gemm_grouped_tn(H1, dU, seg, ids, nb, aligned=rb.aligned, out=o[2])
        gb2 = K.segment_colsum(dU, seg, ids, nb, out=o[3])
        if K.mlp_bwd_grouped_supported(Xg, H1, dU, rb.aligned):
            # dA1 = (dU W2^T) . [H1 > 0] generated inside the grouped product: never written to / re-read from HBM
            gW1, gb1 = K.mlp_bwd_grouped(Xg, H1, dU, W2, seg, ids, nb, out=None if out is None else (o[0], o[1]))
        else:
            dA1 = K.gemm(dU, W2, tb=True, mask=H1)
            gW1 = K.gemm_grouped_tn(Xg, dA1, seg, ids, nb, aligned=rb.aligned, out=o[0])
            gb1 = K.segment_colsum(dA1, seg, ids, nb, out=o[1])
        return [gW1, gb1, gW2, gb2]

    def syn_forward(self, X, A):
        K = self.K
        W1, b1, W2, b2 = self.W
        self.X, self.A = X, A
        self.H1 = K.gemm(X, W1, bias=b1, relu=True)
        U = K.gemm(self.H1, W2, bias=b2)
        self.Tz = [U]
        for _ in range(self.k):
            self.Tz.append(self._prop(A, self.Tz[-1]))
        self.Z = self.Tz[-1]
        return self.Z

    def syn_grads(self):
        K, lay = self.K, self.lay
        N, C, h, nc = self.X.shape[0], self.C, self.h, lay.nblk
        self.S, self.R = K.softmax_residual(self.Z, lay.labels, lay.inv_nc_row)
        self.dTz = [None] * (self.k + 1)
        self.dTz[self.k] = K.expand_class_blocks(self.R, lay.blk, nc)
        for l in range(self.k, 0, -1):
            self.dTz[l - 1] = self._prop(self.A, self.dTz[l], transpose=True)
        dUc = self.dTz[0]                                         # (N, nc*C)
        ones = _ones_col(K, N)
        gW2 = K.gemm(self.H1, dUc, ta=True)
        gb2 = K.colsum(dUc)
        self.dA1c = K.relu_mask(K.gemm(dUc.view(N * nc, C), self.W[2], tb=True), self.H1, groups=nc).view(N, nc * h)
        gW1 = K.gemm(self.X, self.dA1c, ta=True)
        gb1 = K.colsum(self.dA1c)
        return [gW1, gb1, gW2, gb2]

    def syn_backward(self, G, need_dA=True):
        K, lay = self.K, self.lay
        W1, b1, W2, b2 = self.W
        G1, g1, G2, g2 = G
        N, C, h, nc = self.X.shape[0], self.C, self.h, lay.nblk
        ones = _ones_col(K, N)
        # tangent forward along G
        At = K.gemm(self.X, G1)
        K.bias_act(At, g1.view(-1), relu=False)
        Ht = K.relu_mask(At, self.H1, groups=nc)
        Ut = K.gemm(Ht.view(N * nc, h), W2).view(N, nc * C)
        K.gemm(self.H1, G2, out=Ut, beta=1.0)
        K.bias_act(Ut, g2.view(-1), relu=False)
        Tt = [Ut]
        for _ in range(self.k):
            Tt.append(self._prop(self.A, Tt[-1]))
        q = K.pick_class_blocks(Tt[-1], lay.blk, nc)
        dZ = K.softmax_jvp(self.S, q, lay.inv_nc_row)
        # reverse through the tangent network (its cotangents are the first-order quantities)
        dA = None
        if need_dA:
            for l in range(1, self.k + 1):
                dA = self._acc_dA(dA, self.dTz[l], Tt[l - 1])
        dH1 = K.gemm(self.dTz[0], G2, tb=True)
        dX = K.gemm(self.dA1c, G1, tb=True)
        # reverse through the primal network from the softmax-Jacobian term
        dT = dZ
        for l in range(self.k, 0, -1):
            if need_dA:
                dA = self._acc_dA(dA, dT, self.Tz[l - 1])
            dT = self._prop(self.A, dT, transpose=True)
        K.gemm(dT, W2, tb=True, out=dH1, beta=1.0)
        dA1 = K.relu_mask(dH1, self.H1)
        K.gemm(dA1, W1, tb=True, out=dX, beta=1.0)
        return dX, dA

    def train_grads(self, X, A):
        K, lay = self.K, self.lay
        W1, b1, W2, b2 = self.W
        Z = self.syn_forward(X, A)
        _, R = K.softmax_residual(Z, lay.labels, lay.inv_n_row)
        dU = R
        for _ in range(self.k):
            dU = self._prop(A, dU, transpose=True)
        ones = _ones_col(K, X.shape[0])
        gW2 = K.gemm(self.H1, dU, ta=True)
        gb2 = K.colsum(dU).view(-1)
        dA1 = K.gemm(dU, W2, tb=True, mask=self.H1)
        gW1 = K.gemm(X, dA1, ta=True)
        gb1 = K.colsum(dA1).view(-1)
        return [gW1, gb1, gW2, gb2]


# ======================================================================================== GCN, 2 layers
class GCN2(_ModelBase):
    """Z = A (relu(A (X W1) + b1) W2) + b2   (models/gcn.py:8-23, base.py:51-78, layers.py:36-51)."""
    param_shapes = property(lambda s: [(s.d, s.h), (s.h,), (s.h, s.C), (s.C,)])
    widths = property(lambda s: [s.h, s.h, s.C, s.C])
    is_bias = [False, True, False, True]

    def real_grads(self, rb, X_full, ones_full, X_padded=None, out=None):
        K = self.K
        W1, b1, W2, b2 = self.W
        o = out if out is not None else [None] * 4
        outer, inner = rb.blocks_fwd
        Xp = X_full if X_padded is None else X_padded
        T2 = K.spmm(outer.with_global_cols(), Xp)[:, :X_full.shape[1]]   # (A2 X[n_id]) W1 == A2 (X[n_id] W1)
        H1 = K.gemm(T2, W1, bias=b1, relu=True)
        M2 = K.gemm(H1, W2)
        Z = K.bias_act(K.spmm(inner.csr, M2), b2, relu=False)
        _, R = K.softmax_residual(Z, rb.labels, rb.inv_b)
        ids, nb = rb.out_block, self.lay.nblk
        gb2 = K.segment_colsum(R, rb.seg[0], ids, nb, out=o[3])
        dM2 = K.spmm(inner.csr_t, R)
        seg1 = rb.seg[1]
        gW2 = K.gemm_grouped_tn(H1, dM2, seg1, ids, nb, aligned=rb.aligned, out=o[2])
        if K.mlp_bwd_grouped_supported(T2, H1, dM2, rb.aligned):
            gW1, gb1 = K.mlp_bwd_grouped(T2, H1, dM2, W2, seg1, ids, nb, out=None if out is None else (o[0], o[1]))
        else:
            dA1 = K.gemm(dM2, W2, tb=True, mask=H1)
            gb1 = K.segment_colsum(dA1, seg1, ids, nb, out=o[1])
            gW1 = K.gemm_grouped_tn(T2, dA1, seg1, ids, nb, aligned=rb.aligned, out=o[0])
        return [gW1, gb1, gW2, gb2]

    def syn_forward(self, X, A):
        K = self.K
        W1, b1, W2, b2 = self.W
        self.X, self.A = X, A
        self.M1 = K.gemm(X, W1)
        self.H1 = self._prop(A, self.M1, bias=b1, relu=True)
        self.M2 = K.gemm(self.H1, W2)
        self.Z = self._prop(A, self.M2, bias=b2)
        return self.Z

    def syn_grads(self):
        K, lay = self.K, self.lay
        N, C, h, nc = self.X.shape[0], self.C, self.h, lay.nblk
        self.S, self.R = K.softmax_residual(self.Z, lay.labels, lay.inv_nc_row)
        self.Rexp = K.expand_class_blocks(self.R, lay.blk, nc)
        ones = _ones_col(K, N)
        gb2 = K.colsum(self.Rexp)
        self.dM2c = self._prop(self.A, self.Rexp, transpose=True)
        gW2 = K.gemm(self.H1, self.dM2c, ta=True)
        self.dA1c = K.relu_mask(K.gemm(self.dM2c.view(N * nc, C), self.W[2], tb=True), self.H1, groups=nc).view(
            N, nc * h)
        gb1 = K.colsum(self.dA1c)
        self.dM1c = self._prop(self.A, self.dA1c, transpose=True)
        gW1 = K.gemm(self.X, self.dM1c, ta=True)
        return [gW1, gb1, gW2, gb2]

    def syn_backward(self, G, need_dA=True):
        K, lay = self.K, self.lay
        W1, b1, W2, b2 = self.W
        G1, g1, G2, g2 = G
        N, C, h, nc = self.X.shape[0], self.C, self.h, lay.nblk
        ones = _ones_col(K, N)
        M1t = K.gemm(self.X, G1)
        A1t = self._prop(self.A, M1t, fresh=True)
        K.bias_act(A1t, g1.view(-1), relu=False)
        H1t = K.relu_mask(A1t, self.H1, groups=nc)
        M2t = K.gemm(H1t.view(N * nc, h), W2).view(N, nc * C)
        K.gemm(self.H1, G2, out=M2t, beta=1.0)
        Zt = self._prop(self.A, M2t, fresh=True)
        K.bias_act(Zt, g2.view(-1), relu=False)
        q = K.pick_class_blocks(Zt, lay.blk, nc)
        dZ = K.softmax_jvp(self.S, q, lay.inv_nc_row)
        dA = None
        if need_dA:
            dA = self._acc_dA(dA, self.Rexp, M2t)
            dA = self._acc_dA(dA, self.dA1c, M1t)
        dH1 = K.gemm(self.dM2c, G2, tb=True)
        dX = K.gemm(self.dM1c, G1, tb=True)
        dM2 = self._prop(self.A, dZ, transpose=True)
        if need_dA:
            dA = self._acc_dA(dA, dZ, self.M2)
        K.gemm(dM2, W2, tb=True, out=dH1, beta=1.0)
        dA1 = K.relu_mask(dH1, self.H1)
        dM1 = self._prop(self.A, dA1, transpose=True)
        if need_dA:
            dA = self._acc_dA(dA, dA1, self.M1)
        K.gemm(dM1, W1, tb=True, out=dX, beta=1.0)
        return dX, dA

    def train_grads(self, X, A):
        K, lay = self.K, self.lay
        W1, b1, W2, b2 = self.W
        Z = self.syn_forward(X, A)
        _, R = K.softmax_residual(Z, lay.labels, lay.inv_n_row)
        ones = _ones_col(K, X.shape[0])
        gb2 = K.colsum(R).view(-1)
        dM2 = self._prop(A, R, transpose=True)
        gW2 = K.gemm(self.H1, dM2, ta=True)
        dA1 = K.gemm(dM2, W2, tb=True, mask=self.H1)
        gb1 = K.colsum(dA1).view(-1)
        dM1 = self._prop(A, dA1, transpose=True)
        gW1 = K.gemm(X, dM1, ta=True)
        return [gW1, gb1, gW2, gb2]


def build_model(K, kind, d, hidden, n_class, nlayers, ntrans, layout, identity_adj=False):
    """condense_model flag -> engine (reference: eval(args.condense_model)(...), gcond.py:38)."""
    if kind == "SGC":
        if ntrans == 1:
            return SGC1(K, d, hidden, n_class, nlayers, layout, identity_adj)
        if ntrans == 2:
            return SGC2(K, d, hidden, n_class, nlayers, layout, identity_adj)
        raise NotImplementedError("condense_model SGC is implemented for ntrans in {1, 2} (all GCond JSON configs)")
    if kind == "GCN":
        if nlayers != 2:
            raise NotImplementedError("condense_model GCN is implemented for nlayers == 2 (the reference default)")
        return GCN2(K, d, hidden, n_class, nlayers, layout, identity_adj)
    raise NotImplementedError(f"condense_model {kind!r}: this path implements SGC | GCN (BASELINE north star)")
