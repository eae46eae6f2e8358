"""ShardedGCond over NCCL (one process per GPU) against the single-GPU run and the reference fixture: classes AND the
PGE pair rows dealt to two ranks reproduce the un-sharded losses step for step, and the replicas end bit-identical.
Needs >= 2 visible GPUs: skipped (with a message) on a single-GPU lease; when it runs it leaves
gpurun_out/nccl_sharded_gcond.json as a record."""
import json
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from graphslim_b200 import data as gdata
    from graphslim_b200 import parallel
    from graphslim_b200.reduction import create_reducer
    from tests import helpers

    def run(sharded):
        args = helpers.case_args(name, device=f"cuda:{rank}", save_init=False, progress=False, gemm_precision=1)
        args.epochs = 2
        args.device = f"cuda:{rank}"
        raw = helpers.case_graph(name)
        data = gdata.TransAndInd(raw, args.dataset, args.pre_norm)
        helpers.seed_everything(args.seed)
        agent = (parallel.SHARDED[args.method](args.setting, data, args) if sharded
                 else create_reducer(args.method, setting=args.setting, data=data, args=args))
        losses = []
        agent.trace = lambda kind, **kw: losses.append(float(kw["loss"].item())) if kind == "grads" else None
        agent.reduce(data, verbose=False)
        torch.cuda.synchronize()
        return agent, np.array(losses)

    single, l_single = run(False)
    shard, l_shard = run(True)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), l_single=l_single, l_shard=l_shard,
             feat=shard.feat_syn.cpu().numpy(), feat_single=single.feat_syn.cpu().numpy(),
             pge=np.concatenate([p.cpu().numpy().ravel() for p in shard.pge.parameters()]),
             owned=np.array(shard.owned_classes), pge_sharded=np.array(bool(shard.pge_sharded)))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["mini_sgc2_arxiv", "mini_gcn_flickr"])
def test_sharded_gcond_over_nccl_matches_single_gpu(name, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("ShardedGCond over NCCL needs >= 2 visible GPUs (single-GPU lease)")
    world = 2
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    from tests import helpers
    gold = helpers.golden(name)
    r0, r1 = ranks
    assert bool(r0["pge_sharded"]) and bool(r1["pge_sharded"])
    assert set(r0["owned"]).isdisjoint(set(r1["owned"]))
    assert np.array_equal(r0["feat"], r1["feat"]) and np.array_equal(r0["pge"], r1["pge"])        # replicas identical
    np.testing.assert_array_equal(r0["l_shard"], r1["l_shard"])
    n = len(r0["l_shard"])
    tol = helpers.parity_tol(name, 1)
    # the sharded run against the single-GPU run of the same process, and both against the reference fixture
    np.testing.assert_allclose(r0["l_shard"][:2], r0["l_single"][:2], rtol=tol["first_tol"])
    np.testing.assert_allclose(r0["l_shard"], r0["l_single"], rtol=tol["traj_tol"])
    np.testing.assert_allclose(r0["l_shard"][:2], gold["losses"][:2], rtol=tol["first_tol"])
    np.testing.assert_allclose(r0["l_shard"], gold["losses"][:n], rtol=tol["traj_tol"])
    rel = float(np.linalg.norm(r0["feat"] - r0["feat_single"]) / np.linalg.norm(r0["feat_single"]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "nccl_sharded_gcond.json")
    rec = json.load(open(path)) if os.path.exists(path) else {}
    rec[name] = dict(world=world, steps=int(n), max_loss_rel_vs_single=float(np.abs(r0["l_shard"] / r0["l_single"] - 1).max()),
                     max_loss_rel_vs_fixture=float(np.abs(r0["l_shard"] / gold["losses"][:n] - 1).max()),
                     feat_rel_vs_single=rel)
    json.dump(rec, open(path, "w"), indent=1)
