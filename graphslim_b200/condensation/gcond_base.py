"""GCondBase: the shared machinery of GCond / GCondX on the B200 path.

Host-side mirror of graphslim/condensation/gcond_base.py (same constructor contract, same attributes the
callers touch: ``feat_syn, pge, adj_syn, labels_syn, num_class_dict, syn_class_indices, nnodes_syn``), with the
per-class matching loop (:156-241) replaced by the batched closed form in ``graphslim_b200.engine``.
"""
import os
import time
from collections import Counter

import numpy as np
import scipy.sparse as sp
import torch

from .. import engine as _engine
from ..ops import CudaOps, Csr
from ..pge import PGE
from ..sampler import ClassSampler, DeviceClassSampler


def _kernels(device, args):
    """The only compute backend: CUDA kernels from libgraphslim_b200.so (raises on CPU / missing library).
    Optional switches (default on): args.pge_fused -- the fused tcgen05 PGE pipeline (csrc/pge_fused.cu);
    args.grouped_mn -- the TMA-fed MN-major grouped products of the real side (csrc/grouped_tn.cu; they flush a class's
    partial sums from several CTAs with float atomics, so two runs differ in the last bits)."""
    K = CudaOps(device, precision=int(getattr(args, "gemm_precision", 1)))
    K.pge_fused = bool(getattr(args, "pge_fused", True))
    K.grouped_mn = bool(getattr(args, "grouped_mn", True))
    return K


class _Adam:
    """torch.optim.Adam(params, lr) with default betas/eps, one fused kernel per tensor (gcond_base.py:68-69)."""

    def __init__(self, K, params, lr):
        self.K, self.params, self.lr = K, params, float(lr)
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0

    def step(self, grads):
        self.t += 1
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            if g is None:
                continue
            self.K.adam_step(p, g.contiguous().view_as(p), m, v, self.t, self.lr)


class InnerLoop:
    """The condense-model training steps of gcond.py:63-72 (``optimizer_model`` is a fresh Adam every epoch, :44).

    Every step has the same shapes and, with Adam's step-dependent scalars read from a device table
    (gs_adam_step_table_f32), the same launch arguments, so one step is captured in a CUDA graph the first time it
    runs and replayed afterwards: ~15 launches of small kernels become one graph launch (the loop is launch-bound: 78 %
    of a Cora-shape epoch).  The model weights, the optimiser state and the adjacency live in fixed buffers; a new
    epoch copies the fresh weight draw into them and clears the state.  Results are bit-identical to the step-by-step
    path (same kernels, same order, same scalars); `use_graph=False` (or a failed capture) runs exactly that path.
    """

    def __init__(self, K, model, feat_syn, n_syn, lr, steps_per_epoch, use_graph=True, use_chain=False):
        self.K, self.model, self.feat = K, model, feat_syn
        self.W = [K.zeros(*shape) for shape in model.param_shapes]
        self.m = [torch.zeros_like(w) for w in self.W]
        self.v = [torch.zeros_like(w) for w in self.W]
        self.adj = K.zeros(n_syn, n_syn)
        self.table = K.adam_table(steps_per_epoch, float(lr))
        self.capacity = int(steps_per_epoch)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=K.device)
        self.steps_done = 0
        self.use_graph = bool(use_graph) and torch.device(K.device).type == "cuda"
        self.graph = None
        self.warm = False
        self.replays = 0
        # opt-in (args.inner_chain / GS_INNER_CHAIN=1): the whole step as ONE persistent kernel (graphslim_b200/chain.py,
        # csrc/chain.cu) instead of a graph of ~35 launches; exact-fp32 products.  Correct (tests/test_chain_gpu.py) but
        # measured SLOWER than the graph replay -- 0.30 vs 0.18 ms per step at the arxiv shape, 130 vs 77 us at Cora: with
        # one 32 x 32 tile per CTA the small-N products occupy a fifth of the SMs and each operation pays two dependent
        # L2 round trips plus a grid barrier.  Not with a fixed identity adjacency.
        self.use_chain = bool(use_chain) and self.use_graph and not getattr(model, "identity_adj", False)
        self.chain = None

    def begin_epoch(self, W_host):
        """model.initialize() + a fresh Adam: new weights into the fixed buffers, optimiser state cleared."""
        for dst, src in zip(self.W, W_host):
            dst.copy_(src.view_as(dst))
        for t in self.m + self.v:
            t.zero_()
        self.step_dev.zero_()
        self.steps_done = 0
        self.model.set_weights(self.W)
        return self.W

    def set_adj(self, adj):
        if adj.data_ptr() != self.adj.data_ptr():
            self.adj.copy_(adj)

    def _one_step(self):
        K = self.K
        grads = self.model.train_grads(self.feat, self.adj)
        for p, g, m, v in zip(self.W, grads, self.m, self.v):
            K.adam_step_table(p, g.contiguous().view_as(p), m, v, self.table, self.step_dev)
        K.counter_add(self.step_dev, 1)

    def step(self):
        if self.steps_done >= self.capacity:
            raise RuntimeError("InnerLoop: more steps than the Adam table was sized for")
        self.steps_done += 1
        if self.chain is not None:
            self.chain.run()
            self.replays += 1
            return
        if self.graph is not None:
            self.graph.replay()
            self.replays += 1
            return
        if not self.use_graph or not self.warm:
            self._one_step()               # first step runs eagerly: lazily configured kernels, workspace growth
            self.warm = True
            return
        if self.use_chain:
            from .. import chain as _chain
            try:
                self.chain = _chain.record(self.K, [self, self.model], self._one_step)
                self.chain.run()
                self.replays += 1
                return
            except Exception as exc:       # a call the recorder does not know, no cooperative launch: CUDA-graph path
                self.use_chain, self.chain, self.chain_error = False, None, repr(exc)
        graph = torch.cuda.CUDAGraph()
        try:
            _capture(self.K, graph, self._one_step)
        except Exception as exc:           # stay correct on anything the capture cannot express
            self.use_graph = False
            self.capture_error = repr(exc)
            torch.cuda.synchronize(self.K.device)
            self._one_step()
            return
        self.graph = graph
        self.graph.replay()                # capture does not execute: run the step it recorded
        self.replays += 1


def _capture(K, graph, body):
    """Capture `body()` into `graph` on a side stream (capture_begin/capture_end directly: the torch.cuda.graph()
    context manager also empties the caching allocator, which would hand the ~GB PGE activations back to the driver in
    the middle of a run).  Returns body's result; raises whatever the capture raised."""
    side = torch.cuda.Stream(K.device)
    side.wait_stream(torch.cuda.current_stream(K.device))
    with torch.cuda.stream(side):
        graph.capture_begin(capture_error_mode="thread_local")
        try:
            out = body()
        finally:
            graph.capture_end()
    torch.cuda.current_stream(K.device).wait_stream(side)
    return out


class MatchGraph:
    """The fixed-shape middle of an outer step -- synthetic forward, first-order gradients, matching distance and the
    second-order backward (gcond_base.py:210-239 for all classes at once, ~90 small launches) -- captured once in two
    CUDA graphs and replayed.  The first half (forward + first-order gradients) does not depend on the real side, so
    `start` replays it on a side stream while the caller samples and computes the real-side gradients (a host-bound
    sequence of variable-shape launches) on the main stream; `finish` joins the streams and replays the second half
    (matching + second-order backward).  Inputs that are fresh tensors every step (the real-side class-column
    gradients, the normalised adjacency) are copied into fixed buffers first; loss, dX and dA come back in the graph's
    own fixed outputs, which the caller consumes before the next replay.  The side stream's products pack their B
    operand into a workspace of their own (K._ws is only safe in stream order).  First call eager (lazy kernel
    configuration, workspace growth), second call captures, later calls replay; a failed capture falls back to the
    step-by-step path for good."""

    def __init__(self, K, model, feat_syn, metric, use_graph=True, overlap=True, adj_buffer=None):
        self.K, self.model, self.feat, self.metric = K, model, feat_syn, metric
        # adj_buffer: a fixed tensor the caller normally passes as `adj` (the inner loop's adjacency buffer, which the
        # PGE inference writes in place): captured by address, so the steady-state step copies nothing in
        self.adj_buffer = adj_buffer
        self.use_graph = bool(use_graph) and torch.device(K.device).type == "cuda"
        self.overlap = bool(overlap)
        self.graph, self.graph_fwd, self.calls, self.replays = None, None, 0, 0
        self.gr = self.adj = self.out = self.side = self._pending = None

    def _fwd(self, adj):
        self.model.syn_forward(self.feat, adj)
        self._gs = self.model.syn_grads()

    def _bwd(self, gr, need_dA):
        K, model = self.K, self.model
        loss = K.zeros(1)
        G = K.match(self._gs, gr, model.widths, model.is_bias, model.lay.coeff, self.metric, loss)
        dX, dA = model.syn_backward(G, need_dA=need_dA)
        return loss, dX, dA

    def _body(self, gr, adj, need_dA):
        self._fwd(adj)
        return self._bwd(gr, need_dA)

    def start(self, adj, need_dA):
        """Begin the step's synthetic forward (side stream) if the graphs exist; otherwise only remember the inputs."""
        self.calls += 1
        self._pending = (adj, need_dA, False)
        if not self.use_graph or self.graph is None:
            return
        if adj.data_ptr() != self.adj.data_ptr():
            self.adj.copy_(adj)
        if not self.overlap:
            return
        self.side.wait_stream(torch.cuda.current_stream(self.K.device))
        with torch.cuda.stream(self.side):
            self.graph_fwd.replay()
        self._pending = (adj, need_dA, True)

    def finish(self, gr):
        K = self.K
        adj, need_dA, started = self._pending
        self._pending = None
        if not self.use_graph or self.calls == 1:
            return self._body(gr, adj, need_dA)
        if self.graph is None:
            self.gr = [g.clone() for g in gr]
            if self.adj_buffer is not None and self.adj_buffer.shape == adj.shape:
                self.adj = self.adj_buffer
                if adj.data_ptr() != self.adj.data_ptr():
                    self.adj.copy_(adj)
            else:
                self.adj = adj.clone()
            g_fwd, g_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            ws_main = getattr(K, "_ws", None)
            try:
                K._ws = torch.empty(ws_main.numel() if ws_main is not None else 1 << 22, dtype=torch.uint8,
                                    device=K.device)
                try:
                    _capture(K, g_fwd, lambda: self._fwd(self.adj))
                finally:
                    self._side_ws, K._ws = K._ws, ws_main
                self.out = _capture(K, g_bwd, lambda: self._bwd(self.gr, need_dA))
            except Exception as exc:
                self.use_graph, self.capture_error = False, repr(exc)
                torch.cuda.synchronize(K.device)
                return self._body(gr, adj, need_dA)
            self.graph_fwd, self.graph = g_fwd, g_bwd
            self.side = torch.cuda.Stream(K.device)
        else:
            for dst, src in zip(self.gr, gr):
                if src.data_ptr() != dst.data_ptr():       # real_grads(out=...) already wrote into the fixed buffers
                    dst.copy_(src)
        if started:
            torch.cuda.current_stream(K.device).wait_stream(self.side)
        else:
            self.graph_fwd.replay()
        self.graph.replay()
        self.replays += 1
        return self.out

    def real_out(self):
        """The fixed real-side gradient buffers once the graphs exist (hand them to model.real_grads(out=...))."""
        return self.gr if (self.use_graph and self.graph is not None) else None

    def run(self, gr, adj, need_dA):
        self.start(adj, need_dA)
        return self.finish(gr)


class GCondBase:
    def __init__(self, setting, data, args, **kwargs):
        self.data, self.args, self.setting = data, args, setting
        self.device = args.device
        self.K = K = _kernels(self.device, args)
        if args.with_bn or args.dropout != 0 or getattr(args, "multi_label", False) or getattr(args, "soft_label", 0):
            raise NotImplementedError("the B200 GCond path covers with_bn=False, dropout=0, hard single labels "
                                      "(every GCond/GCondX JSON config of the reference)")
        self.labels_syn = self.data.labels_syn = self.generate_labels_syn(data)
        n = self.nnodes_syn = self.data.labels_syn.shape[0]
        self.d = d = data.feat_train.shape[1]
        print(f'target reduced size:{int(data.feat_train.shape[0] * args.reduction_rate)}')
        print(f'actual reduced size:{n}')
        self.feat_syn = torch.empty(n, d, dtype=torch.float32, device=K.device)
        self.pge = PGE(K, nfeat=d, nnodes=n, args=args)
        self.adj_syn = None
        self.optimizer_feat = _Adam(K, [self.feat_syn], args.lr_feat)
        self.optimizer_pge = _Adam(K, self.pge.parameters(), args.lr_adj)
        print('adj_syn:', (n, n), 'feat_syn:', self.feat_syn.shape)
        self.trace = None          # optional callback(kind, **payload) used by the parity tests

    # ------------------------------------------------------------------ gcond_base.py:79-115
    def generate_labels_syn(self, data):
        counter = Counter(data.labels_train.tolist())
        num_class_dict = {}
        n = len(data.labels_train)
        ordered = sorted(counter.items(), key=lambda kv: kv[1])     # stable: ties keep first-seen order
        used, labels_syn = 0, []
        self.syn_class_indices = {}
        for ix, (c, num) in enumerate(ordered):
            if ix == len(ordered) - 1:
                num_class_dict[c] = max(int(n * self.args.reduction_rate) - used, 1)
            else:
                num_class_dict[c] = max(int(num * self.args.reduction_rate), 1)
                used += num_class_dict[c]
            self.syn_class_indices[c] = [len(labels_syn), len(labels_syn) + num_class_dict[c]]
            labels_syn += [c] * num_class_dict[c]
        self.data.num_class_dict = self.num_class_dict = num_class_dict
        if self.args.verbose:
            print(num_class_dict)
        return np.array(labels_syn)

    # ------------------------------------------------------------------ gcond_base.py:117-151 (init='random')
    def init(self, with_adj=False, reuse_init=False):
        """GCondBase.init (gcond_base.py:117-151): Random.select (sparsification/random.py:9-17) + MFCoreSet.reduce
        (model_free_coreset_base.py:16-61), including the `--agg` variant whose initial features are the two-hop
        aggregation A_hat^2 X of the selected nodes (full-graph SpMM, :18-27) and `reuse_init` (:134-143)."""
        args, data = self.args, self.data
        if args.init != "random":
            raise NotImplementedError("init reducers other than 'random' (KCenter / Herding train a GCN on the full "
                                      "graph first) are outside the GCond hot path")
        if reuse_init:
            base = f"{args.save_path}/reduced_graph/{args.init}"
            tag = f"{args.dataset}_{args.reduction_rate}_{args.seed}.pt"
            if os.path.exists(f"{base}/feat_{tag}") and (not with_adj or os.path.exists(f"{base}/adj_{tag}")):
                feat = torch.load(f"{base}/feat_{tag}", map_location="cpu")
                if not with_adj:
                    return feat
                adj = torch.load(f"{base}/adj_{tag}", map_location="cpu")
                return feat, adj
        lt = np.asarray(data.labels_train)
        base = np.arange(len(data.idx_train)) if args.setting == "ind" else np.asarray(data.idx_train)
        picks = [np.random.permutation(base[lt == c])[:cnt] for c, cnt in self.num_class_dict.items()]
        idx = np.hstack(picks)
        self.init_ids = idx
        src_feat = data.feat_full if args.setting == "trans" else data.feat_train
        src_adj = data.adj_full if args.setting == "trans" else data.adj_train
        src_lab = data.labels_full if args.setting == "trans" else data.labels_train
        if getattr(args, "agg", False):
            if args.setting != "trans":
                raise NotImplementedError("--agg init in the inductive setting: the reference multiplies a train-sized "
                                          "operator with feat_full (model_free_coreset_base.py:37-41) and fails")
            K = self.K
            if getattr(self, "adj_csr", None) is None:
                self._prepare_real()
            csr = self._spmm_csr()
            agg = K.spmm(csr, K.spmm(csr, self.features.contiguous()))        # A_hat (A_hat X) == (A_hat A_hat) X
            data.feat_syn = agg[torch.from_numpy(idx).to(K.device)].float().cpu()
            data.adj_syn = torch.eye(len(idx))
        else:
            # the reference keeps the induced adjacency as a coalesced sparse COO tensor (to_tensor of a scipy matrix,
            # utils.py:240-247) -- also the format of the file it saves under reduced_graph/random
            sub = src_adj[np.ix_(idx, idx)].tocoo()
            data.adj_syn = torch.sparse_coo_tensor(torch.from_numpy(np.vstack([sub.row, sub.col]).astype(np.int64)),
                                                   torch.from_numpy(sub.data.astype(np.float32)),
                                                   torch.Size(sub.shape)).coalesce()
            data.feat_syn = torch.as_tensor(src_feat)[torch.from_numpy(idx)].float()
        data.labels_syn = torch.as_tensor(src_lab)[torch.from_numpy(idx)].long()
        if getattr(args, "save_init", True):
            from ..dataset_utils import save_reduced
            keep = args.method
            args.method = args.init
            try:
                save_reduced(data.adj_syn, data.feat_syn, data.labels_syn, args)
            finally:
                args.method = keep
        return (data.feat_syn, data.adj_syn) if with_adj else data.feat_syn

    def _spmm_csr(self):
        """The full-graph normalised CSR with its long-row work items (power-law hubs), for the wide SpMM."""
        if getattr(self, "_adj_csr_chunked", None) is None:
            csr = self.adj_csr
            if torch.device(self.K.device).type == "cuda":
                from ..graph_utils import build_row_chunks, chunks_to_device
                csr = Csr(csr.rowptr, csr.col, csr.val, csr.n_rows, csr.n_cols,
                          chunks_to_device(build_row_chunks(self.adj_host[0]), self.K.device))
            self._adj_csr_chunked = csr
        return self._adj_csr_chunked

    # ------------------------------------------------------------------ real graph in HBM
    def _prepare_real(self):
        """to_tensor + normalize_adj_tensor(sparse=True) (utils.py:220-247,403-413,451-458) and the class lists of
        retrieve_class_sampler (loader.py:188-195)."""
        args, data, K = self.args, self.data, self.K
        if args.setting == 'trans':
            feats, adj, labels = data.feat_full, data.adj_full, data.labels_full
        else:
            feats, adj, labels = data.feat_train, data.adj_train, data.labels_train
        a = sp.csr_matrix(adj, dtype=np.float32)
        a = (a + sp.eye(a.shape[0], dtype=np.float32, format="csr")).tocsr()   # A + I; existing diagonal entries sum
        a.sum_duplicates()
        a.sort_indices()
        n = a.shape[0]
        # degrees and D^-1/2 in float64 on the host (n values), the nnz-sized products on the device
        rowsum = np.asarray(a.astype(np.float64).sum(1)).ravel()
        with np.errstate(divide="ignore"):
            r = np.power(rowsum, -0.5)
        r[np.isinf(r)] = 0.0
        rowptr = torch.from_numpy(a.indptr.astype(np.int32)).to(K.device)
        col = torch.from_numpy(a.indices.astype(np.int32)).to(K.device)
        raw = torch.from_numpy(a.data.astype(np.float32)).to(K.device)
        val = K.csr_gcn_norm(rowptr, col, raw, torch.from_numpy(r).to(K.device))
        # the reference keeps the CSR of the transpose (SparseTensor(...).t()); identical for a symmetric graph
        at = a.T.tocsr()
        at.sort_indices()
        symmetric = (at.indptr.shape == a.indptr.shape and np.array_equal(at.indptr, a.indptr)
                     and np.array_equal(at.indices, a.indices) and np.array_equal(at.data, a.data))
        val_host = val.cpu().numpy()
        if not symmetric:
            m = sp.csr_matrix((val_host, a.indices, a.indptr), shape=a.shape).T.tocsr()
            m.sort_indices()
            a_indptr, a_indices, val_host = m.indptr, m.indices, m.data.astype(np.float32)
            rowptr = torch.from_numpy(a_indptr.astype(np.int32)).to(K.device)
            col = torch.from_numpy(a_indices.astype(np.int32)).to(K.device)
            val = torch.from_numpy(val_host).to(K.device)
        else:
            a_indptr, a_indices = a.indptr, a.indices
        self.adj_csr = Csr(rowptr, col, val, n, n)
        self.adj_host = (a_indptr.astype(np.int64), a_indices.astype(np.int32), val_host)
        # features, padded to a 32-byte multiple per row so feature-row gathers are float4 aligned
        d = feats.shape[1]
        ld = (d + 7) // 8 * 8
        fx = torch.zeros(n, ld, dtype=torch.float32, device=K.device)
        fx[:, :d] = torch.as_tensor(feats).float().to(K.device)
        self.features = fx[:, :d]
        self.features_padded = fx          # zero pad columns: lets the feature-width SpMM use float4 for any d
        self.ones_full = torch.ones(n, 1, dtype=torch.float32, device=K.device)
        lab = np.asarray(labels).astype(np.int32)
        lt = np.asarray(data.labels_train)
        members = []
        for c in range(data.nclass):
            members.append(np.asarray(data.idx_train)[lt == c] if args.setting == 'trans'
                           else np.arange(len(lt))[lt == c])
        # neighbour sampling runs on the device against the resident CSR (bit-identical to the host sampler, which
        # stays available as args.sampler = "host" and is what the CPU-emulated tests use)
        kind = getattr(args, "sampler", "device")
        if kind == "device" and K.device.type == "cuda":
            self.sampler = DeviceClassSampler(self.adj_csr, members, args.dataset, args.nlayers, K.device,
                                              labels=torch.from_numpy(lab))
        else:
            self.sampler = ClassSampler(*self.adj_host, members, args.dataset, args.nlayers, K.device)
            self.sampler.set_labels(lab)
        if self.trace:
            self.trace("norm", rowptr=a_indptr, col=a_indices, val=val_host)

    # ------------------------------------------------------------------ one matching step (gcond_base.py:156-241)
    def match_step(self, model):
        """Returns (loss device scalar, dX, dA_hat, batch) for the current feat_syn / adj_syn / model weights.
        With class sharding (model.lay.mask) only the owned classes are sampled in full and matched."""
        K = self.K
        mg = getattr(self, "match_graph", None)
        use_mg = mg is not None and mg.use_graph and self.trace is None
        if use_mg:
            mg.start(self.adj_syn, not model.identity_adj)    # synthetic forward on a side stream, under the real side
        pf = getattr(self, "_prefetch", None)
        t0 = time.perf_counter()
        with K.timed("phase_sample_h2d"):
            rb = pf.next() if pf is not None else self.sampler.sample(model.lay.mask)
        self.host_wait_sampler_s = getattr(self, "host_wait_sampler_s", 0.0) + time.perf_counter() - t0
        if self.trace:
            self.trace("sample", rb=rb)
        with K.timed("phase_real_grads"):
            gr = model.real_grads(rb, self.features, self.ones_full, self.features_padded,
                                  out=mg.real_out() if use_mg else None)
        if use_mg:
            with K.timed("phase_syn_graph"):                 # (forward + gradients) | matching + backward: two replays
                loss, dX, dA = mg.finish(gr)
            return loss, dX, dA, rb
        with K.timed("phase_syn_forward_grads"):
            model.syn_forward(self.feat_syn, self.adj_syn)
            gs = model.syn_grads()
        with K.timed("phase_match"):
            loss = K.zeros(1)
            G = K.match(gs, gr, model.widths, model.is_bias, model.lay.coeff, self.args.dis_metric, loss)
        with K.timed("phase_syn_backward"):
            dX, dA = model.syn_backward(G, need_dA=not model.identity_adj)
        return loss, dX, dA, rb

    def draw_model_weights(self, model):
        """reset_parameters of MyLinear / GraphConvolution (layers.py:30-34,369-373): weight and bias
        U(-1/sqrt(in), 1/sqrt(in)), drawn with torch's CPU generator in parameter order."""
        W = []
        shapes = model.param_shapes
        for i in range(0, len(shapes), 2):
            fin, fout = shapes[i]
            s = 1.0 / np.sqrt(fin)
            w = torch.empty(fin, fout).uniform_(-s, s)
            b = torch.zeros(fout).uniform_(-s, s)
            W += [w, b]
        return W

    def get_loops(self, args):
        return args.outer_loop, args.inner_loop

    def check_bn(self, model):
        return model          # with_bn is rejected in __init__ (gcond_base.py:261-285 is a no-op without BN)

    # ------------------------------------------------------------------ checkpoint hook (gcond_base.py:287-324)
    def intermediate_evaluation(self, best_val, loss_avg=None, save=True):
        """gcond_base.py:287-324: `run_inter_eval` runs of test_with_val on the published condensed graph; the graph is
        saved when the mean validation accuracy improves.  ``args.evaluator(data, args) -> (val, test)`` replaces the
        built-in GCN evaluator (graphslim_b200/evaluation.py) when given."""
        from ..dataset_utils import save_reduced
        data, args = self.data, self.args
        if args.verbose:
            print('loss_avg: {}'.format(loss_avg))
        evaluator = getattr(args, "evaluator", None)
        if evaluator is None:
            evaluator = lambda d, a: self.test_with_val(setting=a.setting, iters=a.eval_epochs)
        res = np.array([evaluator(data, args) for _ in range(args.run_inter_eval)]).T
        current_val = res[0].mean()
        args.logger.info('\nVal:  {:.4f} +/- {:.4f}'.format(100 * current_val, 100 * res[0].std()))
        args.logger.info('Test: {:.4f} +/- {:.4f}'.format(100 * res[1].mean(), 100 * res[1].std()))
        if save and current_val > best_val:
            best_val = current_val
            save_reduced(data.adj_syn, data.feat_syn, data.labels_syn, args)
        return best_val

    def test_with_val(self, verbose=False, setting='trans', iters=200, best_val=None):
        """gcond_base.py:326-358 -> [validation accuracy, test accuracy] of a fresh eval GCN trained on the condensed
        graph (full-graph validation forward every iteration)."""
        from ..evaluation import GCNEvaluator
        if getattr(self, "_evaluator", None) is None:
            self._evaluator = GCNEvaluator(self.K, self.data, self.args)
        return self._evaluator.test_with_val(iters=iters, setting=setting)
