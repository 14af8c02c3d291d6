"""Timing of the rollout kernel (config 4: 4096 cube tosses x 80 steps), CUDA events."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dair_pll_b200 import ops, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=4096)
ap.add_argument('--steps', type=int, default=80)
ap.add_argument('--reps', type=int, default=5)
a = ap.parse_args()
dev = torch.device('cuda', 0)
system = bench.make_system(dev, torch.float64)
inertia, mu, half = (t.detach() for t in system._cube_params(torch.float64))
x0 = synthetic.cube_states(a.batch, seed=0, device=dev)
for _ in range(2):
    traj, _ = ops.cube_rollout(x0, inertia, mu, half, bench.DT, a.steps)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(a.reps):
    traj, _ = ops.cube_rollout(x0, inertia, mu, half, bench.DT, a.steps)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / a.reps
print(f'rollout B={a.batch} steps={a.steps}: {ms:.3f} ms  {a.batch * a.steps / ms / 1e3:.1f} M steps/s  '
      f'final z mean {traj[:, -1, 6].mean().item():.6f}')
