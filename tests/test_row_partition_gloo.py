"""Row-partitioned full-graph SpMM on CPU: world-size-2 gloo run of the host logic
(nnz-balanced row blocks, padded index space, slab-pipelined all-gather, transposed-block backward + reduce-scatter)
with the local kernel replaced by a scipy CSR product.  The GPU version of this test lives in test_kernels_gpu.py."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _graph(n=700, F=23, seed=3):
    rng = np.random.default_rng(seed)
    deg = np.minimum((rng.pareto(1.2, n) * 3 + 1).astype(np.int64), n // 2)     # power-law rows, a few hubs
    deg[::97] = 0                                                                  # and some empty rows
    rows = np.repeat(np.arange(n), deg)
    cols = rng.integers(0, n, rows.size)
    A = sp.csr_matrix((rng.standard_normal(rows.size).astype(np.float32), (rows, cols)), shape=(n, n))
    A.sum_duplicates()
    A.sort_indices()
    X = rng.standard_normal((n, F)).astype(np.float32)
    return A, X


def _cpu_spmm(csr, X, out):
    out.copy_(torch.from_numpy(np.asarray(csr @ X.numpy(), dtype=np.float32)))
    return out


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from graphslim_b200.parallel import RowPartitionedSpmm
    A, X = _graph()
    op = RowPartitionedSpmm(A.indptr, A.indices, A.data, rank=rank, world=world, device="cpu",
                            make_csr=lambda m: m, spmm=_cpu_spmm, n_slabs=3)
    Xt = torch.from_numpy(X)
    Y = op.forward(op.shard(Xt).contiguous())
    dX = op.backward(op.shard(Xt).contiguous())
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), Y=Y.numpy(), dX=dX.numpy(), lo=op.lo, hi=op.hi,
             nnz=op.nnz_local)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_partitioned_spmm_matches_single_process(world, tmp_path):
    port = 31500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    A, X = _graph()
    Y_ref = np.asarray(A @ X, dtype=np.float32)
    dX_ref = np.asarray(A.T.tocsr() @ X, dtype=np.float32)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    assert parts[0]["lo"] == 0 and parts[-1]["hi"] == A.shape[0]
    for a, b in zip(parts[:-1], parts[1:]):
        assert a["hi"] == b["lo"]
    Y = np.concatenate([p["Y"] for p in parts])
    dX = np.concatenate([p["dX"] for p in parts])
    # every output row is produced by exactly one rank in CSR order: bit-identical to the unpartitioned product
    np.testing.assert_array_equal(Y, Y_ref)
    np.testing.assert_allclose(dX, dX_ref, rtol=1e-5, atol=1e-5)
    nnz = np.array([int(p["nnz"]) for p in parts])
    assert nnz.sum() == A.nnz


def test_partition_rows_by_nnz_balances_work():
    from graphslim_b200.parallel import feature_slabs, partition_rows_by_nnz
    A, _ = _graph(n=5000)
    for world in (1, 2, 4, 8):
        b = partition_rows_by_nnz(A.indptr, world)
        assert b[0] == 0 and b[-1] == 5000 and np.all(np.diff(b) >= 0) and len(b) == world + 1
        work = np.array([A.indptr[b[r + 1]] - A.indptr[b[r]] + (b[r + 1] - b[r]) for r in range(world)])
        assert work.max() <= work.sum() / world + np.diff(A.indptr).max() + 1
    assert partition_rows_by_nnz(np.zeros(1, dtype=np.int64), 2).tolist() == [0, 0, 0]
    for F, s in [(602, 4), (128, 4), (7, 4), (3, 2), (500, 3)]:
        slabs = feature_slabs(F, s)
        assert slabs[0][0] == 0 and slabs[-1][1] == F
        assert all(a[1] == b[0] for a, b in zip(slabs[:-1], slabs[1:])) and all(c0 % 4 == 0 for c0, _ in slabs)
