"""GCond on B200: drop-in for graphslim.condensation.gcond.GCond (same ctor / ``reduce(data, verbose)``)."""
import torch
from tqdm import trange

from .. import engine as _engine
from ..dataset_utils import verbose_time_memory
from .gcond_base import GCondBase, _Adam


class GCond(GCondBase):
    """"Graph Condensation for Graph Neural Networks" -- loop of graphslim/condensation/gcond.py:17-81."""

    x_variant = False

    def __init__(self, setting, data, args, **kwargs):
        super().__init__(setting, data, args, **kwargs)

    def _pge_turn(self, it, ol):
        return it % 50 < 10                                   # gcond.py:58-61

    @verbose_time_memory
    def reduce(self, data, verbose=True):
        args, K, pge = self.args, self.K, self.pge
        labels_syn = torch.as_tensor(data.labels_syn).long()
        self._prepare_real()
        feat_init = self.init()
        self.feat_syn.copy_(feat_init.to(K.device))
        outer_loop, inner_loop = self.get_loops(args)
        loss_avg, best_val = 0, 0
        layout = _engine.ClassLayout(K, self.labels_syn, data.nclass)
        model = _engine.build_model(K, args.condense_model, self.d, args.hidden, data.nclass, args.nlayers,
                                    args.ntrans, layout, identity_adj=self.x_variant)
        self.draw_model_weights(model)                        # constructor draw (model = SGC(...), gcond.py:38)
        n = self.nnodes_syn
        if self.x_variant:
            self.adj_syn = torch.eye(n, device=K.device)      # gcondx.py:32
        mask = getattr(self, "class_mask", None)              # class sharding (parallel.py); None = all classes
        loss_dev = K.zeros(1)
        adj_syn_inner = None
        for it in trange(args.epochs, disable=not getattr(args, "progress", True)):
            W = [w.to(K.device) for w in self.draw_model_weights(model)]      # model.initialize(), gcond.py:42
            model.set_weights(W)
            if self.trace:
                self.trace("model_init", epoch=it, W=W)
            optimizer_model = _Adam(K, W, args.lr)
            loss_dev.zero_()
            for ol in range(outer_loop):
                if not self.x_variant:
                    adj_raw = pge.forward(self.feat_syn)
                    self.adj_syn, r_norm = K.dense_gcn_norm(adj_raw)
                loss, dX, dA, rb = self.match_step(model, mask)
                K.axpby(1.0, loss, 1.0, loss_dev)
                if self.x_variant:
                    pge_grads, feat_grad = None, dX
                else:
                    dA_raw = K.dense_gcn_norm_bwd(dA, self.adj_syn, r_norm)
                    pge_grads, dX_pge = pge.backward(dA_raw)
                    feat_grad = K.axpby(1.0, dX_pge, 1.0, dX)
                feat_grad, pge_grads = self.reduce_grads(feat_grad, pge_grads)
                if self.trace:
                    self.trace("grads", step=(it, ol), loss=loss, feat_grad=feat_grad, pge_grads=pge_grads)
                if self._pge_turn(it, ol):
                    if pge_grads is not None:
                        self.optimizer_pge.step(pge_grads)
                else:
                    self.optimizer_feat.step([feat_grad])
                if self.x_variant:
                    adj_inner = self.adj_syn
                else:
                    adj_syn_inner = pge.inference(self.feat_syn)
                    adj_inner, _ = K.dense_gcn_norm(adj_syn_inner)
                for _ in range(inner_loop):
                    optimizer_model.step(model.train_grads(self.feat_syn, adj_inner))
            self.loss_avg_dev = loss_dev
            if it in args.checkpoints:
                loss_avg = (loss_avg + float(loss_dev.item())) / (data.nclass * outer_loop)
                self._publish(data, adj_syn_inner, labels_syn)
                best_val = self.intermediate_evaluation(best_val, loss_avg)
        self._publish(data, adj_syn_inner, labels_syn)
        return data

    def reduce_grads(self, feat_grad, pge_grads):
        """Hook for class sharding: all-reduce of [feat_syn.grad || PGE grads] (graphslim_b200/parallel.py)."""
        return feat_grad, pge_grads

    def _publish(self, data, adj_syn_inner, labels_syn):
        n = self.nnodes_syn
        if self.x_variant or adj_syn_inner is None:
            self.adj_out = torch.eye(n)                       # gcondx.py:75-76
        else:
            self.adj_out = adj_syn_inner.detach()
        data.adj_syn, data.feat_syn, data.labels_syn = self.adj_out, self.feat_syn.detach(), labels_syn.detach()
