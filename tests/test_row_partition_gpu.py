"""Row-partitioned A_hat X on CUDA through the C-ABI SpMM kernel: NCCL, one process per GPU.  World size = the
number of visible GPUs capped at 2 (a 1-GPU box still exercises the slab pipeline, padding and both collectives)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.test_row_partition_gloo import _graph

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from graphslim_b200.graph_utils import build_row_chunks, chunks_to_device
    from graphslim_b200.ops import Csr, CudaOps
    from graphslim_b200.parallel import RowPartitionedSpmm
    K = CudaOps(f"cuda:{rank}")
    A, X = _graph(n=20000, F=602, seed=11)
    op = RowPartitionedSpmm(A.indptr, A.indices, A.data, rank=rank, world=world, device=K.device, spmm=K.spmm,
                            n_slabs=4, long_row_nnz=256)
    Xd = torch.from_numpy(X).to(K.device)
    Y = op.forward(op.shard(Xd).contiguous())
    dX = op.backward(op.shard(Xd).contiguous())
    # single-GPU kernel on the whole graph (same chunking threshold -> same summation order for short rows)
    mk = lambda a, dt: torch.from_numpy(np.asarray(a).astype(dt)).to(K.device)
    full = Csr(mk(A.indptr, np.int32), mk(A.indices, np.int32), mk(A.data, np.float32), A.shape[0], A.shape[1],
               chunks_to_device(build_row_chunks(A.indptr, 256), K.device))
    Y1 = K.spmm(full, Xd)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), Y=Y.cpu().numpy(), dX=dX.cpu().numpy(),
             Y1=Y1[op.lo:op.hi].cpu().numpy(), deg=np.diff(A.indptr)[op.lo:op.hi])
    dist.destroy_process_group()


def test_row_partitioned_spmm_cuda(tmp_path):
    world = min(2, torch.cuda.device_count())
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    A, X = _graph(n=20000, F=602, seed=11)
    Y_ref = (A.astype(np.float64) @ X.astype(np.float64))
    dX_ref = (A.T.tocsr().astype(np.float64) @ X.astype(np.float64))
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    Y = np.concatenate([p["Y"] for p in parts])
    dX = np.concatenate([p["dX"] for p in parts])
    scale = np.abs(Y_ref).max()
    np.testing.assert_allclose(Y, Y_ref, rtol=1e-4, atol=1e-5 * scale)
    np.testing.assert_allclose(dX, dX_ref, rtol=1e-4, atol=1e-5 * np.abs(dX_ref).max())
    for p in parts:                     # rows that are not split into atomically accumulated slices: bit-identical
        short = p["deg"] <= 256
        np.testing.assert_array_equal(p["Y"][short], p["Y1"][short])
