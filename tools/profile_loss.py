"""Runs the cube loss kernel a few times on a synthetic batch -- the command profiled under ncu
(see profiles/README.md).  Not a benchmark: numbers printed under a profiler are not bench values."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dair_pll_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=1 << 20)
ap.add_argument('--reps', type=int, default=4)
ap.add_argument('--dtype', default='f64')
ap.add_argument('--variant', type=int, default=0)
a = ap.parse_args()
ops.set_loss_variant(a.variant)
dev = torch.device('cuda', 0)
dtype = torch.float64 if a.dtype == 'f64' else torch.float32
system = bench.make_system(dev, dtype)
x, xp = bench.make_batch(system, a.batch, 0, dev, dtype)
inertia, mu, half = (t.detach() for t in system._cube_params(dtype))
for _ in range(a.reps):
    loss, grad, s, _, it = ops.cube_loss_raw(x, xp, inertia, mu, half, bench.DT, 1e-3, want_iters=True)
torch.cuda.synchronize()
print('loss sum', s.item(), 'mean iters', it.double().mean().item(), 'max iters', it.max().item(),
      'hist', torch.bincount(it)[:40].tolist())
