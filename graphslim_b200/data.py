"""Data-object contract of the GCond path: a mirror of the reference's ``TransAndInd``
(graphslim/dataset/loader.py:100-135) that needs neither torch_sparse nor PyG.

The reducers only *read* ``adj_full / feat_full / labels_full / idx_* / adj_train / feat_train /
labels_train / nclass`` and *write* ``labels_syn / num_class_dict / adj_syn / feat_syn``, so an object built by
the reference's own loader works just as well.
"""
import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F


def edge_index_to_csr(edge_index, num_nodes):
    """dataset/convertor.py:71-75 (duplicates are summed by tocsr)."""
    ei = edge_index.numpy() if isinstance(edge_index, torch.Tensor) else np.asarray(edge_index)
    return sp.coo_matrix((np.ones_like(ei[0]), (ei[0], ei[1])), shape=(num_nodes, num_nodes)).tocsr()


class TransAndInd:
    def __init__(self, data, dataset, norm=True):
        self.num_nodes = int(data.num_nodes)
        self.train_mask, self.val_mask, self.test_mask = data.train_mask, data.val_mask, data.test_mask
        self.x, self.y = data.x, data.y
        self.feat_full, self.labels_full = data.x, data.y
        self.adj_full = edge_index_to_csr(data.edge_index, self.num_nodes)
        self.edge_index = data.edge_index
        if dataset in ("flickr", "reddit", "ogbn-arxiv"):
            from sklearn.preprocessing import StandardScaler     # loader.py:113-119
            # torch tensors go to sklearn exactly as the reference passes them: check_array turns them into float64, so
            # the standardisation is computed in double and rounded to fp32 once (a float32 ndarray would be
            # transformed in place in fp32 and differ from the reference in the last bit)
            scaler = StandardScaler()
            scaler.fit(self.x[data.idx_train])
            self.feat_full = torch.from_numpy(scaler.transform(self.x)).float()
        if norm and dataset in ("cora", "citeseer", "pubmed"):
            self.feat_full = F.normalize(self.feat_full, p=1, dim=1)
        self.idx_train, self.idx_val, self.idx_test = data.idx_train, data.idx_val, data.idx_test
        it, iv, ite = (np.asarray(t) for t in (self.idx_train, self.idx_val, self.idx_test))
        self.adj_train = self.adj_full[np.ix_(it, it)]
        self.adj_val = self.adj_full[np.ix_(iv, iv)]
        self.adj_test = self.adj_full[np.ix_(ite, ite)]
        self.labels_train = self.labels_full[self.idx_train]
        self.labels_val = self.labels_full[self.idx_val]
        self.labels_test = self.labels_full[self.idx_test]
        self.feat_train = self.feat_full[self.idx_train]
        self.feat_val = self.feat_full[self.idx_val]
        self.feat_test = self.feat_full[self.idx_test]
        self.nclass = int(getattr(data, "num_classes", int(self.labels_full.max()) + 1))
        self.labels_syn, self.feat_syn, self.adj_syn = None, None, None
