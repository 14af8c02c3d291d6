"""Kernel-only and whole-step timings of the cube loss for natural / cost-ordered batches and static / dynamic chunk
scheduling, over batch sizes (A/B builds: DAIR_PLL_B200_LIB=... python tools/exp_order.py).  Not a benchmark."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dair_pll_b200 import ops, parallel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--sizes', default='65536,131072,262144,1048576')
ap.add_argument('--no-step', action='store_true')
a = ap.parse_args()
dev = torch.device('cuda', 0)
system = bench.make_system(dev, torch.float64)
X, XP = bench.make_batch(system, 1 << 20, 0, dev, torch.float64)
lt, ct = system.multibody_terms.lagrangian_terms, system.multibody_terms.contact_terms
leaves = [t.detach() for t in (lt.inertial_parameters, ct.friction_params, ct.geometries[0].length_params)]
params = list(system.parameters())
tag = os.environ.get('DAIR_PLL_B200_LIB', 'default')


def t_ms(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


for B in [int(v) for v in a.sizes.split(',')]:
    x, xp = X[:B].contiguous(), XP[:B].contiguous()
    it = ops.cube_loss_leaf_dp_raw(x, xp, *leaves, bench.DT, 1e-3, want_iters=True)[4]
    order = torch.argsort(it, descending=True, stable=True)
    xo, xpo = x[order].contiguous(), xp[order].contiguous()
    reps = max(20, int(40.0 / (0.1 + 0.45 * B / 2 ** 20)))
    row = {}
    for name, (xx, xxp) in {'natural': (x, xp), 'cost': (xo, xpo)}.items():
        for fname, flags in {'static': 0, 'dynamic': ops.LOSS_DYNAMIC}.items():
            row[f'{name}/{fname}'] = t_ms(lambda: ops.cube_loss_leaf_dp_raw(xx, xxp, *leaves, bench.DT, 1e-3, flags=flags,
                                                                            want_loss=True), reps)
    msg = f'[{tag}] B={B}: kernel(3 launches) ms ' + '  '.join(f'{k} {v:.4f}' for k, v in row.items())
    if not a.no_step:
        for name, (xx, xxp, dyn) in {'natural/static': (x, xp, False), 'cost/dynamic': (xo, xpo, True)}.items():
            system.dynamic_schedule = dyn

            def step():
                for p in params:
                    p.grad = None
                m = system.contactnets_loss(xx, None, xxp).mean()
                m.backward()
                return m
            g = parallel.GraphedStep(step, dev)
            msg += f'  | step[{name}] graph {t_ms(g, reps):.4f} eager {t_ms(step, 20):.4f}'
    print(msg, flush=True)
