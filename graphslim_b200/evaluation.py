"""Checkpoint evaluator of the GCond path: the reference trains a fresh 2-layer GCN on the
condensed graph and validates it on the real graph after EVERY training iteration.

    GCondBase.intermediate_evaluation  graphslim/condensation/gcond_base.py:287-324
    GCondBase.test_with_val            :326-358
    BaseGNN.fit_with_val / test / predict   graphslim/models/base.py:80-225   (GCN, mode='eval': dropout 0, wd 5e-4)

Training runs on the dense synthetic graph through the same closed-form GCN gradients the inner loop uses
(engine.GCN2.train_grads); every validation / test forward is a full-graph propagation A_hat (X W) -- the wide
CSR SpMM (gs_spmm_csr_f32, width = hidden) followed by the class-width one -- which is where the standalone SpMM
kernel earns its keep inside a condensation run.  No autograd, no torch_sparse.
"""
import numpy as np
import scipy.sparse as sp
import torch

from . import engine as _engine
from .ops import Csr


def normalized_csr(K, adj):
    """normalize_adj_tensor(adj, sparse=True) (graphslim/utils.py:403-413,451-458) of a scipy matrix, resident in HBM:
    D^-1/2 (A + I) D^-1/2 with the float64 products rounded once to fp32 (gs_csr_gcn_norm_f64)."""
    a = sp.csr_matrix(adj, dtype=np.float32)
    a = (a + sp.eye(a.shape[0], dtype=np.float32, format="csr")).tocsr()
    a.sum_duplicates()
    a.sort_indices()
    rowsum = np.asarray(a.astype(np.float64).sum(1)).ravel()
    with np.errstate(divide="ignore"):
        r = np.power(rowsum, -0.5)
    r[np.isinf(r)] = 0.0
    dev = K.device
    rowptr = torch.from_numpy(a.indptr.astype(np.int32)).to(dev)
    col = torch.from_numpy(a.indices.astype(np.int32)).to(dev)
    raw = torch.from_numpy(a.data.astype(np.float32)).to(dev)
    val = K.csr_gcn_norm(rowptr, col, raw, torch.from_numpy(r).to(dev))
    val_host = val.cpu().numpy()
    indptr, indices = a.indptr, a.indices
    nt = sp.csr_matrix((val_host, a.indices, a.indptr), shape=a.shape).T.tocsr()
    nt.sort_indices()
    # the reference multiplies with SparseTensor(...).t(): the transpose of the NORMALISED matrix.  It equals the matrix
    # itself only when structure AND values are symmetric (a directed or weighted graph is not)
    if not (np.array_equal(nt.indptr, indptr) and np.array_equal(nt.indices, indices)
            and np.array_equal(nt.data, val_host)):
        indptr, indices = nt.indptr, nt.indices
        rowptr = torch.from_numpy(indptr.astype(np.int32)).to(dev)
        col = torch.from_numpy(indices.astype(np.int32)).to(dev)
        val = torch.from_numpy(nt.data.astype(np.float32)).to(dev)
    chunks = None
    if dev.type == "cuda":
        # long-row work items are built from the row pointer of the CSR that is actually returned (the transpose of a
        # directed graph has different row lengths than the graph)
        from .graph_utils import build_row_chunks, chunks_to_device
        chunks = chunks_to_device(build_row_chunks(indptr), dev)
    return Csr(rowptr, col, val, a.shape[0], a.shape[1], chunks)


def _draw_gcn_weights(shapes):
    """GraphConvolution.reset_parameters (models/layers.py:30-34): weight and bias U(-1/sqrt(in), 1/sqrt(in)) from
    torch's CPU generator, in parameter order."""
    W = []
    for i in range(0, len(shapes), 2):
        fin, fout = shapes[i]
        s = 1.0 / np.sqrt(fin)
        W += [torch.empty(fin, fout).uniform_(-s, s), torch.zeros(fout).uniform_(-s, s)]
    return W


class _Graph:
    """A real graph the evaluator propagates over: normalised CSR + features in HBM, uploaded once per reducer."""

    def __init__(self, K, adj, feat):
        self.csr = normalized_csr(K, adj)
        self.X = torch.as_tensor(feat).float().to(K.device).contiguous()


class GCNEvaluator:
    def __init__(self, K, data, args):
        self.K, self.data, self.args = K, data, args
        if getattr(args, "eval_model", "GCN") != "GCN" or args.nlayers != 2:
            raise NotImplementedError("the checkpoint evaluator implements eval_model GCN with nlayers == 2 "
                                      "(the reference default)")
        self._graphs = {}

    def _graph(self, which):
        if which not in self._graphs:
            d = self.data
            adj, feat = {"full": (d.adj_full, d.feat_full), "val": (d.adj_val, d.feat_val),
                         "test": (d.adj_test, d.feat_test)}[which]
            self._graphs[which] = _Graph(self.K, adj, feat)
        return self._graphs[which]

    def _predict(self, g, W):
        """argmax of BaseGNN.forward on a real graph (log_softmax is monotone): A (relu(A (X W1) + b1) W2) + b2."""
        K = self.K
        W1, b1, W2, b2 = W
        H1 = K.bias_act(K.spmm(g.csr, K.gemm(g.X, W1)), b1, relu=True)
        Z = K.bias_act(K.spmm(g.csr, K.gemm(H1, W2)), b2, relu=False)
        return Z.argmax(1)

    @staticmethod
    def _accuracy(pred, labels):
        return float((pred == labels).double().mean().item())            # utils.accuracy

    def test_with_val(self, iters=None, setting=None):
        """One run of gcond_base.py:326-358: returns [best validation accuracy, test accuracy of that model]."""
        K, data, args = self.K, self.data, self.args
        iters = int(args.eval_epochs if iters is None else iters)
        setting = args.setting if setting is None else setting
        dev = K.device
        feat = torch.as_tensor(data.feat_syn).float().to(dev).contiguous()
        adj_raw = torch.as_tensor(data.adj_syn).float().to(dev).contiguous()
        labels_syn = np.asarray(torch.as_tensor(data.labels_syn).cpu())
        n_syn, d = feat.shape
        layout = _engine.ClassLayout(K, labels_syn, data.nclass)
        model = _engine.GCN2(K, d, args.hidden, data.nclass, 2, layout)
        _draw_gcn_weights(model.param_shapes)                             # GCN(...) constructor draw
        W = [w.to(dev) for w in _draw_gcn_weights(model.param_shapes)]     # fit_with_val: self.initialize()
        model.set_weights(W)
        self.last_init = [w.clone() for w in W]
        adj, _ = K.dense_gcn_norm(adj_raw)                               # normalize_adj_tensor(adj, sparse=False)
        if setting == "ind":
            g_val, idx_val = self._graph("val"), None
        else:
            g_val, idx_val = self._graph("full"), torch.as_tensor(np.asarray(data.idx_val)).long().to(dev)
        labels_val = torch.as_tensor(np.asarray(data.labels_val)).long().to(dev)
        lr, wd = float(args.lr), 5e-4                                     # BaseGNN mode == 'eval'
        m = [torch.zeros_like(w) for w in W]
        v = [torch.zeros_like(w) for w in W]
        t, best, best_W, curve = 0, 0.0, None, []
        for i in range(iters):
            if i == iters // 2 and lr > 0.001:                            # base.py:146-147: new optimiser, lr / 10
                lr, t = lr * 0.1, 0
                for s in m + v:
                    s.zero_()
            grads = model.train_grads(feat, adj)
            t += 1
            for p, g, mm, vv in zip(W, grads, m, v):
                g = g.contiguous().view_as(p)
                K.axpby(wd, p, 1.0, g)                                    # Adam(weight_decay): g += wd * p
                K.adam_step(p, g, mm, vv, t, lr)
            pred = self._predict(g_val, W)
            acc = self._accuracy(pred if idx_val is None else pred[idx_val], labels_val)
            curve.append(acc)
            if acc > best:
                best, best_W = acc, [w.clone() for w in W]
        if best_W is not None:
            W = best_W                                                    # load_state_dict(weights)
        labels_test = torch.as_tensor(np.asarray(data.labels_test)).long().to(dev)
        if setting == "ind":
            acc_test = self._accuracy(self._predict(self._graph("test"), W), labels_test)
        else:
            idx_test = torch.as_tensor(np.asarray(data.idx_test)).long().to(dev)
            acc_test = self._accuracy(self._predict(self._graph("full"), W)[idx_test], labels_test)
        self.last_val_curve = curve
        return [best, acc_test]
