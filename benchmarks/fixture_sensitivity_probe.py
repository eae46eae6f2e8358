"""How well conditioned is a parity fixture's loss trajectory?  CPU only.

Runs a fixture case with every kernel replaced by its PyTorch reference (tests/emu_ops.py) and multiplies each dense
product by (1 + eps * N(0,1)) drawn from a private generator (the samplers' random streams stay untouched), then prints
the per-step relative deviation of the loss from the reference fixture.  A step whose deviation jumps by orders of
magnitude under eps = 1e-6 sits on a ReLU kink / near-zero column norm: its bound in tests/helpers.py::PARITY_TOL has to
cover that jump whatever kernel computes the products.

    python benchmarks/fixture_sensitivity_probe.py mini_doscond_gcn
    -> step 6 moves by 1.7e-2 in one of two eps = 1e-6 runs, <= 5e-4 everywhere else
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphslim_b200.condensation import gcond_base            # noqa: E402
from tests import test_engine_emulated as T                   # noqa: E402
from tests.emu_ops import EmuOps                              # noqa: E402

GEN = torch.Generator().manual_seed(5)


class NoisyOps(EmuOps):
    eps = 0.0

    def gemm(self, *a, **k):
        out = super().gemm(*a, **k)
        if self.eps:
            out.mul_(1 + self.eps * torch.randn(out.shape, generator=GEN))
        return out


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "mini_doscond_gcn"
    gcond_base._kernels = lambda device, args: NoisyOps(device)
    for eps in (0.0, 1e-7, 1e-6, 1e-6, 3e-6, 3e-6):
        NoisyOps.eps = eps
        gold, sub, args, data, agent, seen, pge_init = T.run_case(name)
        got = np.array(seen["losses"])
        ref = gold["losses"][:len(got)]
        print(f"eps {eps:7.0e}  max {np.max(np.abs(got - ref) / ref):.1e}  per step",
              np.array2string(np.abs(got - ref) / ref, precision=1, max_line_width=200))


if __name__ == "__main__":
    main()
