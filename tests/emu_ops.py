"""TEST INFRASTRUCTURE: plain-PyTorch fp32 reference of every kernel in graphslim_b200.ops.CudaOps.

Two uses:
  * ``-m gpu`` tests compare each CUDA kernel against the method of the same name here;
  * ``-m "not gpu"`` tests run the host-side engine (closed-form matching, PGE backward, loops) on top of these
    references to check the *algebra* against the oracle where no GPU exists.
It is never imported by the product package; the product has no CPU path.
"""
import numpy as np
import torch

from graphslim_b200.ops import Csr


class EmuOps:
    def __init__(self, device="cpu", precision=0, fused=True):
        self.device = torch.device(device)
        self.precision = precision
        self._launches = 0
        self.pge_fused = bool(fused)       # True: the algebra of the fused layer-2 pipeline (the product's default path)

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=torch.float32):
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    def launches(self):
        return self._launches

    def timed(self, tag):
        import contextlib
        return contextlib.nullcontext()

    # ---- dense
    def gemm(self, A, B, ta=False, tb=False, out=None, alpha=1.0, beta=0.0, precision=None, bias=None, relu=False,
             mask=None):
        a = A.T if ta else A
        b = B.T if tb else B
        r = alpha * (a @ b)
        if out is not None and beta != 0.0:
            r = out * beta + r
        if bias is not None:
            r = r + bias
        if relu:
            r = r.clamp(min=0)
        if mask is not None:
            r = r * (mask > 0).to(r.dtype)
        if out is None:
            return r
        out.copy_(r)
        return out

    def gemm_grouped_tn(self, A, B, seg, out_block, nblk, aligned=False, precision=None, out=None):
        M, N = A.shape[1], B.shape[1]
        out = torch.zeros(M, nblk * N, device=self.device) if out is None else out.zero_()
        seg = seg.tolist()
        for g, ob in enumerate(out_block.tolist()):
            r0, r1 = seg[g], seg[g + 1]
            out[:, ob * N:(ob + 1) * N] = A[r0:r1].T @ B[r0:r1]
        return out

    def mlp_bwd_grouped_supported(self, X, H1, dU, aligned):
        return bool(getattr(self, "fuse_mlp_bwd", True))

    def mlp_bwd_grouped(self, X, H1, dU, W2, seg, out_block, nblk, out=None):
        dA1 = (dU @ W2.T) * (H1 > 0)
        o0, o1 = (None, None) if out is None else out
        return (self.gemm_grouped_tn(X, dA1, seg, out_block, nblk, out=o0),
                self.segment_colsum(dA1, seg, out_block, nblk, out=o1))

    def segment_colsum(self, X, seg, out_block, nblk, out=None):
        cols = X.shape[1]
        out = torch.zeros(1, nblk * cols, device=self.device) if out is None else out.zero_()
        sl = seg.tolist()
        for g, ob in enumerate(out_block.tolist()):
            out[0, ob * cols:(ob + 1) * cols] = X[sl[g]:sl[g + 1]].sum(0)
        return out

    def colsum(self, X):
        return X.sum(0, keepdim=True)

    # ---- sparse
    @staticmethod
    def _coo(csr):
        counts = (csr.rowptr[1:] - csr.rowptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(csr.n_rows, device=csr.rowptr.device), counts)
        return rows, csr.col.long()

    def spmm(self, csr, X, out=None, accumulate=False, tile_cols=None):
        rows, cols = self._coo(csr)
        y = torch.zeros(csr.n_rows, X.shape[1], device=self.device)
        y.index_add_(0, rows, csr.val[:, None] * X[cols])
        if out is None:
            return y
        if accumulate:
            out.add_(y)
        else:
            out.copy_(y)
        return out

    def spmm_scatter(self, csr, dY, out):
        rows, cols = self._coo(csr)
        out.index_add_(0, cols, csr.val[:, None] * dY[rows])
        return out

    def gather_rows(self, X, idx):
        return X[idx.long()].contiguous()

    def csr_gcn_norm(self, rowptr, col, a, r64):
        counts = (rowptr[1:] - rowptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(rowptr.numel() - 1, device=rowptr.device), counts)
        return ((r64[rows] * a.double()) * r64[col.long()]).float()

    # ---- glue
    def bias_act(self, Z, bias, relu):
        if bias is not None:
            Z.add_(bias)
        if relu:
            Z.clamp_(min=0)
        return Z

    def relu_mask(self, D, H, groups=1):
        rows, cols = H.shape
        v = D.view(rows, groups, cols)
        v.mul_((H > 0).to(D.dtype)[:, None, :])
        return D

    def softmax_residual(self, Z, labels, row_scale, want_nll=False):
        ls = torch.log_softmax(Z, dim=1)
        S = ls.exp()
        Y = torch.zeros_like(S)
        Y[torch.arange(Z.shape[0]), labels.long()] = 1.0
        sc = row_scale[:, None] if row_scale is not None else 1.0
        R = (S - Y) * sc
        if want_nll:
            return S, R, -ls[torch.arange(Z.shape[0]), labels.long()]
        return S, R

    def expand_class_blocks(self, R, blk, nblk):
        rows, C = R.shape
        E = torch.zeros(rows, nblk, C, device=self.device)
        own = blk >= 0
        E[torch.arange(rows)[own], blk.long()[own]] = R[own]
        return E.view(rows, nblk * C)

    def pick_class_blocks(self, Zf, blk, nblk):
        rows = Zf.shape[0]
        C = Zf.shape[1] // nblk
        out = Zf.view(rows, nblk, C)[torch.arange(rows), blk.long().clamp(min=0)].contiguous()
        out[blk < 0] = 0.0
        return out

    def softmax_jvp(self, S, Q, row_scale):
        q = Q * (row_scale[:, None] if row_scale is not None else 1.0)
        return S * (q - (S * q).sum(1, keepdim=True))

    # ---- matching
    def match(self, gs_list, gr_list, widths, is_bias, coeff, metric, loss_accum):
        nc = coeff.numel()
        out = []
        eps = 1e-6
        if metric == "cos":
            dot = torch.zeros(nc)
            ns2 = torch.zeros(nc)
            nr2 = torch.zeros(nc)
            for a, b, w in zip(gs_list, gr_list, widths):
                av, bv = a.reshape(a.shape[0], nc, w), b.reshape(b.shape[0], nc, w)
                dot += (av * bv).sum((0, 2))
                ns2 += (av * av).sum((0, 2))
                nr2 += (bv * bv).sum((0, 2))
            ns, nr = ns2.sqrt(), nr2.sqrt()
            den = ns * nr + eps
            loss_accum += (coeff * (1 - dot / den)).sum()
            al = -coeff / den
            be = torch.where(ns > 0, coeff * dot * nr / (ns * den * den), torch.zeros_like(ns))
            for a, b, w in zip(gs_list, gr_list, widths):
                A_ = al.repeat_interleave(w)[None, :]
                B_ = be.repeat_interleave(w)[None, :]
                out.append(A_ * b + B_ * a)
            return out
        for a, b, w, bias in zip(gs_list, gr_list, widths, is_bias):
            co = coeff.repeat_interleave(w)
            if metric == "mse":
                loss_accum += (co[None, :] * (a - b) ** 2).sum()
                out.append(2 * co[None, :] * (a - b))
                continue
            if bias:
                out.append(torch.zeros_like(a))
                continue
            dot = (a * b).sum(0)
            ns, nr = a.norm(dim=0), b.norm(dim=0)
            den = ns * nr + eps
            loss_accum += (co * (1 - dot / den)).sum()
            al = -co / den
            be = torch.where(ns > 0, co * dot * nr / (ns * den * den), torch.zeros_like(ns))
            out.append(al[None, :] * b + be[None, :] * a)
        return out

    # ---- dense norm
    def dense_gcn_norm(self, A, out=None):
        n = A.shape[0]
        mx = A + torch.eye(n, device=self.device)
        r = mx.sum(1).pow(-0.5)
        r[torch.isinf(r)] = 0.0
        res = (r[:, None] * mx) * r[None, :]
        return (res if out is None else out.copy_(res)), r

    def dense_gcn_norm_bwd(self, dAhat, Ahat, r):
        rd = (dAhat * Ahat).sum(1)
        cd = (dAhat * Ahat).sum(0)
        drho = -0.5 * r * r * (rd + cd)
        return r[:, None] * r[None, :] * dAhat + drho[:, None]

    # ---- PGE
    @staticmethod
    def _chunks(chunk_off):
        o = chunk_off.tolist()
        return list(zip(o[:-1], o[1:]))

    @staticmethod
    def _y1(Pa, Pb):
        n, h = Pa.shape
        return (Pb[:, None, :] + Pa[None, :, :]).reshape(n * n, h)       # row k = i*n + j -> Pa[j] + Pb[i]

    def _stats(self, Y, chunk_off, eps):
        mean, rstd = [], []
        for a, b in self._chunks(chunk_off):
            y = Y[a:b].double()
            mean.append(y.mean(0))
            rstd.append(1.0 / torch.sqrt(y.var(0, unbiased=False) + eps))
        return torch.stack(mean).float(), torch.stack(rstd).float()

    def _per_row(self, v, chunk_off, rows):
        """(nchunk, h) -> (rows, h) by chunk membership."""
        out = torch.empty(rows, v.shape[1], device=self.device)
        for c, (a, b) in enumerate(self._chunks(chunk_off)):
            out[a:b] = v[c]
        return out

    def pge_l1_stats(self, Pa, Pb, chunk_off, eps=1e-5):
        return self._stats(self._y1(Pa, Pb), chunk_off, eps)

    def pge_l1_stats_closed(self, Pa, Pb, eps=1e-5):
        a, b = Pa.double(), Pb.double()
        var = a.var(0, unbiased=False) + b.var(0, unbiased=False)
        mean = a.mean(0) + b.mean(0)
        return (mean.float().view(1, -1), (1.0 / torch.sqrt(var + eps)).float().view(1, -1),
                torch.stack([a.mean(0), b.mean(0)]).float())

    def pge_bn1_bwd_closed(self, dH1, Pa, Pb, mean, rstd, gamma, beta, col_mean):
        n = Pa.shape[0]
        off = torch.tensor([0, n * n])
        s1, s2 = self.pge_bn1_bwd_stats(dH1, Pa, Pb, off, mean, rstd, gamma, beta)
        dPa, dPb = self.pge_bn1_bwd_reduce(dH1, Pa, Pb, off, mean, rstd, gamma, beta, s1, s2)
        return dPa, dPb, s2.sum(0), s1.sum(0)

    # ---- row-sharded PGE pieces (same contracts as CudaOps)
    def pge_l1_expand_rows(self, Pa, Pb_rows, off_rows, mean, rstd, gamma, beta):
        n, h = Pa.shape
        y = (Pb_rows[:, None, :] + Pa[None, :, :]).reshape(-1, h)
        return torch.relu(gamma * ((y - mean) * rstd) + beta)

    def col_stats_partial(self, Y, off_rows):
        y = Y.double() - Y[0].double()
        return torch.cat([y.sum(0), (y * y).sum(0)])

    def col_stats_combine(self, parts, counts, eps=1e-5):
        h = parts.shape[1] // 3
        n_r = counts.double()[:, None]
        S1, S2, shift = parts[:, :h], parts[:, h:2 * h], parts[:, 2 * h:]
        mu_r = shift + S1 / n_r
        mu = (n_r * mu_r).sum(0) / n_r.sum()
        m2 = ((S2 - S1 * S1 / n_r) + n_r * (mu_r - mu) ** 2).sum(0)
        var = (m2 / n_r.sum()).clamp_(min=0)
        return mu.float().view(1, -1), (1.0 / torch.sqrt(var + eps)).float().view(1, -1)

    def pge_bn1_bwd_pass_rows(self, dH1_rows, Pa, Pb, i_first, n_i, mean, rstd, gamma, beta):
        n, h = Pa.shape
        y = (Pb[i_first:i_first + n_i, None, :] + Pa[None, :, :])                 # (n_i, n, h)
        xh = (y - mean.view(1, 1, h)) * rstd.view(1, 1, h)
        g = dH1_rows.view(n_i, n, h) * ((gamma * xh + beta) > 0)
        work = torch.zeros(2 * h + n * h, dtype=torch.float64)
        work[:h] = g.double().sum((0, 1))
        work[h:2 * h] = (g.double() * xh.double()).sum((0, 1))
        fl = work[2 * h:].view(torch.float32)
        fl[:n * h] = g.sum(0).reshape(-1)                                         # Ga[j]
        Gb = torch.zeros(n, h)
        Gb[i_first:i_first + n_i] = g.sum(1)
        fl[n * h:] = Gb.reshape(-1)
        return work

    def pge_bn1_bwd_final(self, Pa, Pb, rstd, gamma, col_mean, work):
        n, h = Pa.shape
        tsum = work[:2 * h]
        fl = work[2 * h:].view(torch.float32)
        Ga, Gb = fl[:n * h].view(n, h), fl[n * h:].view(n, h)
        m = float(n) * float(n)
        a1n = (tsum[:h] * n / m).float()
        a2 = (tsum[h:] / m).float()
        rs = rstd.view(-1)
        dPa = gamma * rs * (Ga - a1n - rs * n * (Pa - col_mean[0]) * a2)
        dPb = gamma * rs * (Gb - a1n - rs * n * (Pb - col_mean[1]) * a2)
        return dPa, dPb, tsum[h:].float(), tsum[:h].float()

    # ---- fused layer-2 pipeline (same contracts as CudaOps; plain fp32 algebra)
    def pge_fused_supported(self, h, nchunks):
        return bool(self.pge_fused) and int(nchunks) == 1 and int(h) in (128, 256)

    def _h1_rows(self, Pa, Pb, i_first, n_i, bn1):
        mean, rstd, gamma, beta = bn1
        return self.pge_l1_expand_rows(Pa, Pb[i_first:i_first + n_i], None, mean, rstd, gamma, beta)

    def _dy2_rows(self, Y2, dE, bn2, w3, s1, s2, count):
        mean, rstd, gamma, beta = bn2
        xh = (Y2 - mean) * rstd
        d = dE[:, None] * w3[None, :] * ((gamma * xh + beta) > 0)
        return gamma * rstd * (d - s1.view(1, -1) / count - xh * s2.view(1, -1) / count)

    def pge_fused_l2_fwd(self, Pa, Pb, i_first, n_i, mean1, rstd1, gamma1, beta1, W2):
        H1 = self._h1_rows(Pa, Pb, i_first, n_i, (mean1, rstd1, gamma1, beta1))
        Y2 = H1 @ W2.T
        y = Y2.double()
        return Y2, torch.cat([y.sum(0), (y * y).sum(0)])

    def pge_stats_finalize(self, stats, count, eps=1e-5):
        h = stats.numel() // 2
        m = stats[:h] / count
        var = (stats[h:] / count - m * m).clamp_(min=0)
        return m.float().view(1, h), (1.0 / torch.sqrt(var + eps)).float().view(1, h)

    def pge_bn1_work(self, n, h):
        return torch.zeros(2 * h + n * h, dtype=torch.float64)

    def pge_fused_l2_bwd_dx(self, Pa, Pb, i_first, n_i, bn1, W2, Y2, dE, bn2, w3, s1, s2, count, work=None,
                            store=False):
        n, h = Pa.shape
        dH1 = self._dy2_rows(Y2, dE, bn2, w3, s1, s2, count) @ W2
        if store:
            return dH1
        g = (dH1 * (self._h1_rows(Pa, Pb, i_first, n_i, bn1) > 0)).view(n_i, n, h)
        fl = work[2 * h:].view(torch.float32)
        fl[:n * h] += g.sum(0).reshape(-1)
        fl[n * h:].view(n, h)[i_first:i_first + n_i] += g.sum(1)
        return work

    def pge_fused_l2_bwd_dw(self, Pa, Pb, i_first, n_i, bn1, Y2, dE, bn2, w3, s1, s2, count):
        return self._dy2_rows(Y2, dE, bn2, w3, s1, s2, count).T @ self._h1_rows(Pa, Pb, i_first, n_i, bn1)

    def pge_bn1_tsum(self, Pa, Pb, col_mean, rstd1, work):
        n, h = Pa.shape
        fl = work[2 * h:].view(torch.float32)
        Ga, Gb = fl[:n * h].view(n, h).double(), fl[n * h:].view(n, h).double()
        work[:h] += Ga.sum(0)
        work[h:2 * h] += rstd1.view(-1).double() * ((Ga * (Pa.double() - col_mean[0].double())).sum(0)
                                                    + (Gb * (Pb.double() - col_mean[1].double())).sum(0))
        return work

    def pge_l1_expand(self, Pa, Pb, chunk_off, mean, rstd, gamma, beta):
        y = self._y1(Pa, Pb)
        rows = y.shape[0]
        xh = (y - self._per_row(mean, chunk_off, rows)) * self._per_row(rstd, chunk_off, rows)
        return torch.relu(gamma * xh + beta)

    def col_stats_chunked(self, Y, chunk_off, eps=1e-5):
        return self._stats(Y, chunk_off, eps)

    def pge_l3(self, Y2, chunk_off, mean, rstd, gamma, beta, w3, b3):
        rows = Y2.shape[0]
        xh = (Y2 - self._per_row(mean, chunk_off, rows)) * self._per_row(rstd, chunk_off, rows)
        return torch.relu(gamma * xh + beta) @ w3 + b3

    def pge_symm_sigmoid(self, E, n):
        e = E.view(n, n)
        a = torch.sigmoid((e + e.T) / 2)
        return a - torch.diag(torch.diag(a))

    def pge_symm_sigmoid_bwd(self, dA, A):
        n = A.shape[0]
        t = dA * A * (1 - A) * (1 - torch.eye(n, device=self.device))
        return ((t + t.T) / 2).reshape(-1)

    def _l3_common(self, Y2, dE, chunk_off, mean, rstd, gamma, beta, w3):
        rows = Y2.shape[0]
        xh = (Y2 - self._per_row(mean, chunk_off, rows)) * self._per_row(rstd, chunk_off, rows)
        yh = gamma * xh + beta
        dyh = dE[:, None] * w3[None, :] * (yh > 0)
        return xh, yh, dyh

    def pge_l3_bwd_stats(self, Y2, dE, chunk_off, mean, rstd, gamma, beta, w3):
        xh, yh, dyh = self._l3_common(Y2, dE, chunk_off, mean, rstd, gamma, beta, w3)
        s1 = torch.stack([dyh[a:b].sum(0) for a, b in self._chunks(chunk_off)])
        s2 = torch.stack([(dyh[a:b] * xh[a:b]).sum(0) for a, b in self._chunks(chunk_off)])
        dw3 = (dE[:, None] * torch.relu(yh)).sum(0)
        db3 = dE.sum().view(1)
        return s1, s2, dw3, db3

    def pge_bn2_bwd_apply(self, Y2, dE, chunk_off, mean, rstd, gamma, beta, w3, s1, s2):
        rows = Y2.shape[0]
        xh, yh, dyh = self._l3_common(Y2, dE, chunk_off, mean, rstd, gamma, beta, w3)
        m = torch.empty(rows, 1)
        for a, b in self._chunks(chunk_off):
            m[a:b] = float(b - a)
        S1, S2 = self._per_row(s1, chunk_off, rows), self._per_row(s2, chunk_off, rows)
        return gamma * self._per_row(rstd, chunk_off, rows) * (dyh - S1 / m - xh * S2 / m)

    def _bn1_common(self, dH1, Pa, Pb, chunk_off, mean, rstd, gamma, beta):
        y = self._y1(Pa, Pb)
        rows = y.shape[0]
        xh = (y - self._per_row(mean, chunk_off, rows)) * self._per_row(rstd, chunk_off, rows)
        dyh = dH1 * ((gamma * xh + beta) > 0)
        return xh, dyh

    def pge_bn1_bwd_stats(self, dH1, Pa, Pb, chunk_off, mean, rstd, gamma, beta):
        xh, dyh = self._bn1_common(dH1, Pa, Pb, chunk_off, mean, rstd, gamma, beta)
        s1 = torch.stack([dyh[a:b].sum(0) for a, b in self._chunks(chunk_off)])
        s2 = torch.stack([(dyh[a:b] * xh[a:b]).sum(0) for a, b in self._chunks(chunk_off)])
        return s1, s2

    def pge_bn1_bwd_reduce(self, dH1, Pa, Pb, chunk_off, mean, rstd, gamma, beta, s1, s2):
        n, h = Pa.shape
        rows = n * n
        xh, dyh = self._bn1_common(dH1, Pa, Pb, chunk_off, mean, rstd, gamma, beta)
        m = torch.empty(rows, 1)
        for a, b in self._chunks(chunk_off):
            m[a:b] = float(b - a)
        S1, S2 = self._per_row(s1, chunk_off, rows), self._per_row(s2, chunk_off, rows)
        dY1 = gamma * self._per_row(rstd, chunk_off, rows) * (dyh - S1 / m - xh * S2 / m)
        dY1 = dY1.view(n, n, h)                                            # [i, j, :]
        return dY1.sum(0), dY1.sum(1)                                     # dPa[j], dPb[i]

    # ---- optimiser
    def adam_step(self, p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1 = 1 - beta1 ** step
        bc2 = 1 - beta2 ** step
        denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
        p.addcdiv_(m, denom, value=-(lr / bc1))

    def adam_table(self, steps, lr, beta1=0.9, beta2=0.999):
        import numpy as np
        t = np.arange(1, max(int(steps), 1) + 1, dtype=np.float64)
        tab = np.stack([lr / (1.0 - beta1 ** t), np.sqrt(1.0 - beta2 ** t)], 1).astype(np.float32)
        return torch.from_numpy(tab)

    def adam_step_table(self, p, g, m, v, table, step_dev, beta1=0.9, beta2=0.999, eps=1e-8):
        t = int(step_dev.item())
        step_size, bc2_sqrt = float(table[t, 0]), float(table[t, 1])
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)

    def counter_add(self, counter, inc=1):
        counter.add_(int(inc))

    def axpby(self, a, x, b, y):
        if b == 0.0:
            y.copy_(a * x)
        else:
            y.mul_(b).add_(a * x)
        return y
