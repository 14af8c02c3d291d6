"""Timing of the prediction-loss path (rollout + backward) for the cube: reverse-mode adjoint vs forward-mode tangents."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dair_pll_b200 import ops, synthetic  # noqa: E402

dev = torch.device('cuda', 0)
system = bench.make_system(dev, torch.float64)
inertia, mu, half = (t.detach() for t in system._cube_params(torch.float64))
for n, steps in ((4096, 80), (65536, 80), (256, 80)):
    x0 = synthetic.cube_states(n, seed=5, device=dev)
    w = torch.randn(n, steps + 1, 13, device=dev, dtype=torch.float64)
    for fm in (False, True):
        if fm and n > 4096:
            continue
        leaves = [t.clone().requires_grad_() for t in (x0, inertia, mu, half)]

        def step():
            for t in leaves:
                t.grad = None
            traj = ops.CubeRollout.apply(leaves[0], leaves[1], leaves[2], leaves[3], bench.DT, steps, 1e-4, fm)
            (traj * w).sum().backward()
        ms = bench._time_gpu(step, dev, 3, warmup=1)

        def fwd():
            with torch.no_grad():
                ops.cube_rollout(x0, inertia, mu, half, bench.DT, steps)
        ms_f = bench._time_gpu(fwd, dev, 5, warmup=1)
        print(f'{n} x {steps}: forward {ms_f:.3f} ms; forward+backward ({"forward-mode" if fm else "reverse-mode"}) {ms:.3f} ms', flush=True)
