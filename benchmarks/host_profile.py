"""Where does the HOST time of an epoch go?  cProfile of the main thread over a few epochs of the bench workload
(rank 0 prints; works under torchrun for the sharded path).  From N = 4 on the step is bound by the host's launch
rate, so this is the profile that matters for the multi-GPU curve.

    python benchmarks/host_profile.py [--workload ogbn-arxiv] [--epochs 3]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/host_profile.py
"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ogbn-arxiv")
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--top", type=int, default=45)
    ns = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from graphslim_b200.reduction import create_reducer
    raw, args, gdata = bench.make_problem(ns.workload, local, epochs=ns.epochs + 3, gemm_precision=1, track_loss=False)
    data = gdata.TransAndInd(raw, args.dataset, args.pre_norm)
    bench.seed_everything(args.seed)
    if world > 1:
        from graphslim_b200 import parallel
        agent = parallel.SHARDED[args.method](args.setting, data, args)
    else:
        agent = create_reducer(args.method, setting=args.setting, data=data, args=args)
    agent.setup(data)
    for it in range(3):
        agent.run_epoch(it)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    for it in range(3, 3 + ns.epochs):
        agent.run_epoch(it)
    pr.disable()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    if rank == 0:
        outer = agent.get_loops(args)[0]
        print(f"world {world}: host issue {host * 1e3 / ns.epochs / outer:.3f} ms per outer step (profiled), "
              f"with device drain {total * 1e3 / ns.epochs / outer:.3f} ms")
        for key in ("tottime", "cumtime"):
            buf = io.StringIO()
            pstats.Stats(pr, stream=buf).sort_stats(key).print_stats(ns.top)
            print(buf.getvalue()[:9000])
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
