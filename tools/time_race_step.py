"""Public-API step (loss.mean().backward()) of a strong-scaling shard, eager kernel-only vs the CUDA-graph replay, with and
without the racing kernel for the expensive head.  CUDA events; not the bench."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dair_pll_b200 import ops, parallel  # noqa: E402

dev = torch.device('cuda', 0)
system = bench.make_system(dev, torch.float64)
X, XP = bench.make_batch(system, 1 << 20, 0, dev, torch.float64)
lt, ct = system.multibody_terms.lagrangian_terms, system.multibody_terms.contact_terms
leaves = [t.detach() for t in (lt.inertial_parameters, ct.friction_params, ct.geometries[0].length_params)]
hint = ops.cube_loss_leaf_dp_raw(X, XP, *leaves, bench.DT, 1e-3, want_iters=True)[4]
order = torch.argsort(hint, descending=True, stable=True)
params = list(system.parameters())
system.dynamic_schedule = True


def t_us(fn, reps=300):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


for world in (8, 4):
    idx = order[0::world]
    x, xp = X.index_select(0, idx).contiguous(), XP.index_select(0, idx).contiguous()

    def step():
        for p in params:
            p.grad = None
        m = system.contactnets_loss(x, None, xp).mean()
        m.backward()
        return m.detach()
    for race in (False, True):
        system.race_expensive_head = race
        flags = ops.LOSS_DYNAMIC | (ops.LOSS_RACE if race else 0)
        k = t_us(lambda: ops.cube_loss_leaf_dp_raw(x, xp, *leaves, bench.DT, 1e-3, flags=flags))
        g = parallel.GraphedStep(step, dev)
        print(f'shard 1/{world} ({x.shape[0]} pairs) race={race}: entry point {k:.1f} us   graph-replayed step {t_us(g):.1f} us', flush=True)
