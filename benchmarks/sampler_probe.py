#!/usr/bin/env python
"""Times the device sampler alone (one outer step = all classes) on a benchmark-shaped graph.
    python benchmarks/sampler_probe.py [workload] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "reddit"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    from graphslim_b200.reduction import create_reducer
    raw, args, gdata = bench.make_problem(workload, 0, epochs=1, track_loss=False)
    data = gdata.TransAndInd(raw, args.dataset, args.pre_norm)
    bench.seed_everything(args.seed)
    agent = create_reducer(args.method, setting=args.setting, data=data, args=args)
    agent.setup(data)
    s = agent.sampler
    torch.cuda.synchronize()
    for i in range(steps):
        batch, off = s.draw_batches()
        s.checkout_rng()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(s.side):
            a.record(s.side)
        slot = s.launch(batch, off)
        with torch.cuda.stream(s.side):
            b.record(s.side)
        rb = s.collect(slot)
        torch.cuda.synchronize()
        s.release(slot)
        s.checkin_rng()
        print(f"step {i}: device {a.elapsed_time(b):.3f} ms, host {1e3 * (time.perf_counter() - t0):.3f} ms, "
              f"levels {rb.counts}, nnz {[int(b_.csr.col.numel()) for b_ in rb.blocks]}, draws {rb.draws}")


    import numpy as np
    dbg = np.zeros(16, dtype=np.int64)
    s.lib.gs_dsampler_debug_counters(s.handle, dbg.ctypes.data, s.side.cuda_stream)
    if dbg.any():
        print("serial-kernel cycles per step: phase1 %d phase2 %d phase3 %d phase4+reduce %d | whole kernel %d" %
              tuple(int(x / steps) for x in (dbg[0], dbg[1], dbg[2], dbg[3], dbg[8])))


if __name__ == "__main__":
    main()
