"""Per-visit latency of the wavefront loss kernel: a batch made of ONE sample replicated (so every lane of every
warp runs the same Newton chain), for a hard and an easy sample, at one warp / one warp per SM sub-partition /
full residency.  Not a benchmark."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dair_pll_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
system = bench.make_system(dev, torch.float64)
X, XP = bench.make_batch(system, 1 << 17, 0, dev, torch.float64)
lt, ct = system.multibody_terms.lagrangian_terms, system.multibody_terms.contact_terms
leaves = [t.detach() for t in (lt.inertial_parameters, ct.friction_params, ct.geometries[0].length_params)]
it = ops.cube_loss_leaf_dp_raw(X, XP, *leaves, bench.DT, 1e-3, want_iters=True)[4]


def t_us(fn, reps=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


for target in (int(it.max()), 20, 8, 3, 0):
    idx = int((it == target).nonzero()[0])
    for B in (32, 128, 32 * 148 * 4, 32 * 148 * 8, 64 * 148 * 8, 128 * 148 * 8):
        x, xp = X[idx:idx + 1].expand(B, 13).contiguous(), XP[idx:idx + 1].expand(B, 13).contiguous()
        row = []
        for flags in (0, ops.LOSS_DYNAMIC):
            row.append(t_us(lambda: ops.cube_loss_leaf_dp_raw(x, xp, *leaves, bench.DT, 1e-3, flags=flags)))
        print(f'iters={target:2d} B={B:7d} ({B // 32} chunks): static {row[0]:8.1f} us  dynamic {row[1]:8.1f} us', flush=True)
