"""Kernel-only timing of the cube loss kernel (CUDA events), for A/B builds: DAIR_PLL_B200_LIB=... python tools/time_loss.py"""
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
ap.add_argument('--reps', type=int, default=20)
ap.add_argument('--dtype', default='f64')
ap.add_argument('--variant', type=int, default=0)
a = ap.parse_args()
ops.set_loss_variant(a.variant)
dev = torch.device('cuda', 0)
dtype = torch.float64 if a.dtype == 'f64' else torch.float32
system = bench.make_system(dev, dtype)
x, xp = bench.make_batch(system, a.batch, 0, dev, dtype)
inertia, mu, half = (t.detach().to(dtype) for t in system._cube_params(dtype))
for _ in range(3):
    ops.cube_loss_raw(x, xp, inertia, mu, half, bench.DT, 1e-3)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(a.reps):
    out = ops.cube_loss_raw(x, xp, inertia, mu, half, bench.DT, 1e-3)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / a.reps
print(f'{os.environ.get("DAIR_PLL_B200_LIB", "default")}: B={a.batch} {a.dtype} variant={a.variant} '
      f'{ms:.4f} ms/launch  {a.batch / ms / 1e6:.1f} M samples/s  loss_sum={out[2].item():.12e}')
