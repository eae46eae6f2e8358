"""Row-sharded PGE over NCCL (one process per GPU, world = visible GPUs capped at 2): forward adjacency and every
parameter / feature gradient equal the unsharded PGE on the same inputs; the ranks end with identical results."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from types import SimpleNamespace
    from graphslim_b200.ops import CudaOps
    from graphslim_b200.pge import PGE
    K = CudaOps(f"cuda:{rank}", precision=1)
    n, d = 301, 64
    args = SimpleNamespace(dataset="ogbn-arxiv", reduction_rate=0.01)
    torch.manual_seed(5)
    pge = PGE(K, d, n, args)
    x = torch.randn(n, d).to(K.device)
    dA = torch.randn(n, n).to(K.device)
    A_ref = pge.forward(x)
    g_ref, dX_ref = pge.backward(dA)
    sharded = pge.enable_row_sharding(None)
    A = pge.forward(x)
    g, dX = pge.backward(dA)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), sharded=np.array(sharded), A=A.cpu().numpy(), A_ref=A_ref.cpu().numpy(),
             dX=dX.cpu().numpy(), dX_ref=dX_ref.cpu().numpy(),
             **{f"g{i}": t.cpu().numpy() for i, t in enumerate(g)}, **{f"gr{i}": t.cpu().numpy() for i, t in enumerate(g_ref)})
    dist.destroy_process_group()


def test_row_sharded_pge_matches_unsharded(tmp_path):
    world = min(2, torch.cuda.device_count())
    port = 34500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for p in parts:
        assert bool(p["sharded"]) == (world > 1)
        np.testing.assert_allclose(p["A"], p["A_ref"], rtol=1e-4, atol=1e-6)
        scale = np.abs(p["dX_ref"]).max()
        np.testing.assert_allclose(p["dX"], p["dX_ref"], rtol=2e-3, atol=2e-4 * scale)
        for i in range(10):
            ref = p[f"gr{i}"]
            np.testing.assert_allclose(p[f"g{i}"].reshape(ref.shape), ref, rtol=2e-3, atol=2e-4 * (np.abs(ref).max() + 1e-30))
    for p in parts[1:]:                      # replicas: the adjacency bit for bit; gradients up to the summation order
        assert np.array_equal(p["A"], parts[0]["A"])      # of the atomically accumulated split-K products
        for i in range(10):
            ref = parts[0][f"g{i}"]
            np.testing.assert_allclose(p[f"g{i}"], ref, rtol=1e-4, atol=1e-6 * (np.abs(ref).max() + 1e-30))
