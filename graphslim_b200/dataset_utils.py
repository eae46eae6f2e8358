"""On-disk result format and the reduce() decorator of the reference, kept so condensed graphs interoperate
(SURVEY.md section 8f-4).

save_reduced: graphslim/dataset/utils.py:136-152 -- three torch.save files
``{save_path}/reduced_graph/{method}/{adj,feat,label}_{dataset}_{reduction_rate}_{seed}.pt``.
load_reduced / get_syn_data / sparsify: graphslim/dataset/utils.py:8-66,155-296 -- what the reference's evaluator reads
back, including the per-method `threshold` truncation of the learned adjacency.
verbose_time_memory: graphslim/evaluation/utils.py:115-175.
"""
import os
import time
from functools import wraps

import numpy as np
import scipy.sparse as sp
import torch


def save_reduced(adj_syn=None, feat_syn=None, labels_syn=None, args=None):
    save_path = _reduced_dir(args)
    os.makedirs(save_path, exist_ok=True)
    tag = f'{args.dataset}_{args.reduction_rate}_{args.seed}.pt'
    if adj_syn is not None:
        torch.save(adj_syn, os.path.join(save_path, 'adj_' + tag))
    if feat_syn is not None:
        torch.save(feat_syn, os.path.join(save_path, 'feat_' + tag))
    if labels_syn is not None:
        torch.save(labels_syn, os.path.join(save_path, 'label_' + tag))
    args.logger.info(f"Saved {os.path.join(save_path, 'adj_' + tag)}")


def _reduced_dir(args):
    base_path = os.path.abspath(os.path.expanduser(args.save_path))
    save_path = os.path.join(base_path, 'reduced_graph', args.method)
    if getattr(args, "attack", None) is not None and args.dataset in ['flickr']:
        save_path = os.path.join(base_path, 'corrupt_graph', args.attack, 'reduced_graph', args.method)
    return save_path


def _resolve_device(args):
    """dataset/utils.py:162-185: the requested device if it exists, else the nearest one that does."""
    dev = getattr(args, 'device', 'cpu')
    if dev is None or (isinstance(dev, str) and dev.lower() == 'cpu'):
        return 'cpu'
    if isinstance(dev, str) and dev.startswith('cuda'):
        if not torch.cuda.is_available():
            args.logger.warning("CUDA requested but not available. Falling back to CPU for reduced graph loading.")
            return 'cpu'
        try:
            idx = int(dev.split(':')[1]) if ':' in dev else 0
        except ValueError:
            idx = 0
        if idx < torch.cuda.device_count():
            return f'cuda:{idx}'
        args.logger.warning(f"Requested device {dev} unavailable. Using cuda:0 for reduced graph loading.")
        return 'cuda:0'
    return dev


def load_reduced(args, data=None):
    """dataset/utils.py:155-255: (adj_syn, feat_syn, labels_syn) from the three files; every missing piece falls back
    to the original graph's (features / labels of the training part, identity adjacency) exactly as the reference."""
    save_path = _reduced_dir(args)
    target = _resolve_device(args)
    if hasattr(args, 'device') and args.device != target:
        args.device = target
    tag = f'{args.dataset}_{args.reduction_rate}_{args.seed}.pt'

    def load(kind):
        path = os.path.join(save_path, f'{kind}_{tag}')
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        try:
            return torch.load(path, map_location=target)
        except Exception:
            t = torch.load(path, map_location='cpu')
            return t.to(target) if hasattr(t, 'to') and target != 'cpu' else t

    missing = 0
    try:
        feat_syn = load('feat')
    except Exception as e:
        print(f"find no feat at {os.path.join(save_path, 'feat_' + tag)}, use original feature matrix instead. Error: {e}")
        missing += 1
        feat_syn = data.feat_full if args.setting == 'trans' else data.feat_train
    try:
        labels_syn = load('label')
    except Exception as e:
        print(f"find no label at {os.path.join(save_path, 'label_' + tag)}, use original label matrix instead. Error: {e}")
        missing += 1
        labels_syn = data.labels_train
    try:
        adj_syn = load('adj')
    except Exception as e:
        print(f"find no adj at {os.path.join(save_path, 'adj_' + tag)}, use identity matrix instead. Error: {e}")
        missing += 1
        adj_syn = torch.eye(feat_syn.size(0), device=target)
    if missing == 3:
        args.logger.info("no file found, use original graph instead")
    return adj_syn, feat_syn, labels_syn


def sparsify(model_type, adj_syn, args, verbose=False):
    """dataset/utils.py:8-66: the evaluator-side truncation of the learned adjacency -- entries below the per-method
    threshold are zeroed (gcond / doscond: the JSON `threshold` for GNN evaluators, 0.5 / 0.1 for GAT; structure-free
    methods untouched); an MLP evaluator gets the identity."""
    threshold = 0
    if model_type == 'MLP':
        adj_syn = adj_syn - adj_syn
        torch.diagonal(adj_syn).fill_(1)
    elif model_type == 'GAT':
        if args.method in ['gcond', 'doscond']:
            threshold = 0.5 if args.dataset in ['cora', 'citeseer'] else 0.1
        elif args.method in ['msgc']:
            threshold = args.threshold
        else:
            threshold = 0.5
    elif args.method in ['gcond', 'doscond']:
        threshold = args.threshold
    if threshold > 0:
        adj_syn[adj_syn < threshold] = 0
        if verbose:
            print('Sparsity after truncating:', adj_syn.nonzero().shape[0] / adj_syn.numel())
    return adj_syn


def get_syn_data(data, args, model_type, verbose=False):
    """dataset/utils.py:258-296: what an evaluator trains on -- the saved condensed graph, densified and truncated;
    the whole original training graph when nothing was condensed."""
    adj_syn, feat_syn, labels_syn = load_reduced(args, data)
    if labels_syn.shape[0] == data.labels_train.shape[0]:
        return feat_syn, adj_syn, labels_syn
    if isinstance(adj_syn, torch.Tensor) and adj_syn.layout != torch.strided:
        adj_syn = adj_syn.to_dense()
    adj_syn = sparsify(model_type, adj_syn, args, verbose=verbose)
    return feat_syn, adj_syn, labels_syn


def getsize_mb(elements):
    """evaluation/utils.py:42-78 (edge lists counted as 2 x nnz int64 for scipy matrices)."""
    size = 0
    for e in elements:
        if isinstance(e, sp.spmatrix):
            size += 2 * e.nnz * 8
        elif isinstance(e, torch.Tensor):
            size += e.element_size() * e.nelement()
        else:
            t = torch.from_numpy(np.asarray(e))
            size += t.element_size() * t.nelement()
    return size / 1024 / 1024


def verbose_time_memory(func):
    @wraps(func)
    def wrapper(*args, **kwargs):
        verbose = kwargs.get('verbose', False)
        if verbose:
            start = time.perf_counter()
        result = func(*args, **kwargs)
        if verbose:
            run_time = time.perf_counter() - start
            print("Function Time:", run_time, "s")
            print("Function Time:", run_time * 1000, "ms")
            data = kwargs.get('data', None)
            if data is None:
                for arg in args:
                    if hasattr(arg, 'feat_train') or hasattr(arg, 'x'):
                        data = arg
                        break
                if data is None:
                    raise ValueError("The function must be called with 'data' as an argument.")
            origin = getsize_mb([data.feat_train, data.adj_train, data.labels_train])
            condensed = getsize_mb([data.feat_syn, data.adj_syn, data.labels_syn])
            print(f'Original graph:{origin:.2f} Mb  Condensed graph:{condensed:.2f} Mb')
        return result
    return wrapper
