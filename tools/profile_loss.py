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
ap.add_argument('--order', choices=['cost', 'natural'], default='cost')
a = ap.parse_args()
dev = torch.device('cuda', 0)
dtype = torch.float64 if a.dtype == 'f64' else torch.float32
system = bench.make_system(dev, dtype)
x, xp = bench.make_batch(system, a.batch, 0, dev, dtype)
lt, ct = system.multibody_terms.lagrangian_terms, system.multibody_terms.contact_terms
leaves = [t.detach().to(dtype) for t in (lt.inertial_parameters, ct.friction_params, ct.geometries[0].length_params)]
flags = 0
if a.order == 'cost':
    it = ops.cube_loss_leaf_dp_raw(x, xp, *leaves, bench.DT, 1e-3, want_iters=True)[4]
    idx = torch.argsort(it, descending=True, stable=True)
    x, xp = x.index_select(0, idx).contiguous(), xp.index_select(0, idx).contiguous()
    flags = ops.LOSS_DYNAMIC
for _ in range(a.reps):
    loss, sums, means, local, it = ops.cube_loss_leaf_dp_raw(x, xp, *leaves, bench.DT, 1e-3, flags=flags, want_iters=True)
torch.cuda.synchronize()
print('loss sum', sums[15].item(), 'mean iters', it.double().mean().item(), 'max iters', it.max().item(),
      'hist', torch.bincount(it)[:40].tolist())
