"""On-disk result format and the reduce() decorator of the reference, kept so condensed graphs interoperate.

save_reduced: graphslim/dataset/utils.py:136-152 -- three torch.save files
``{save_path}/reduced_graph/{method}/{adj,feat,label}_{dataset}_{reduction_rate}_{seed}.pt``.
verbose_time_memory: graphslim/evaluation/utils.py:115-175.
"""
import os
import time
from functools import wraps

import numpy as np
import scipy.sparse as sp
import torch


def save_reduced(adj_syn=None, feat_syn=None, labels_syn=None, args=None):
    base_path = os.path.abspath(os.path.expanduser(args.save_path))
    save_path = os.path.join(base_path, 'reduced_graph', args.method)
    if getattr(args, "attack", None) is not None and args.dataset in ['flickr']:
        save_path = os.path.join(base_path, 'corrupt_graph', args.attack, 'reduced_graph', args.method)
    os.makedirs(save_path, exist_ok=True)
    tag = f'{args.dataset}_{args.reduction_rate}_{args.seed}.pt'
    if adj_syn is not None:
        torch.save(adj_syn, os.path.join(save_path, 'adj_' + tag))
    if feat_syn is not None:
        torch.save(feat_syn, os.path.join(save_path, 'feat_' + tag))
    if labels_syn is not None:
        torch.save(labels_syn, os.path.join(save_path, 'label_' + tag))
    args.logger.info(f"Saved {os.path.join(save_path, 'adj_' + tag)}")


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
