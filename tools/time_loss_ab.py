"""A/B timing of cube loss kernel builds: DAIR_PLL_B200_LIB=<variant .so> python tools/time_loss_ab.py -- the 1M bench batch in
natural order (static ranges) and in cost order (dynamic chunks), its 1/8 shard with racing warps (the N = 8 strong-scaling
launch) and a 65,536-pair batch; kernel only, CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dair_pll_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
system = bench.make_system(dev, torch.float64)
X, XP = bench.make_batch(system, 1 << 20, 0, dev, torch.float64)
lt, ct = system.multibody_terms.lagrangian_terms, system.multibody_terms.contact_terms
leaves = [t.detach() for t in (lt.inertial_parameters, ct.friction_params, ct.geometries[0].length_params)]


def t_us(fn, reps=100):
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


out = ops.cube_loss_leaf_dp_raw(X, XP, *leaves, bench.DT, 1e-3, want_iters=True)
order = torch.argsort(out[4], descending=True, stable=True)
xo, xpo = X.index_select(0, order).contiguous(), XP.index_select(0, order).contiguous()
idx = order[0::8]
xs, xps = X.index_select(0, idx).contiguous(), XP.index_select(0, idx).contiguous()
nat = t_us(lambda: ops.cube_loss_leaf_dp_raw(X, XP, *leaves, bench.DT, 1e-3))
cost = t_us(lambda: ops.cube_loss_leaf_dp_raw(xo, xpo, *leaves, bench.DT, 1e-3, flags=ops.LOSS_DYNAMIC))
shard = t_us(lambda: ops.cube_loss_leaf_dp_raw(xs, xps, *leaves, bench.DT, 1e-3, flags=ops.LOSS_DYNAMIC | ops.LOSS_RACE), 300)
small = t_us(lambda: ops.cube_loss_leaf_dp_raw(X[:65536], XP[:65536], *leaves, bench.DT, 1e-3), 300)
print(f'{os.path.basename(os.environ.get("DAIR_PLL_B200_LIB", "shipped"))}: 1M natural {nat:.1f} us  1M cost order {cost:.1f} us  '
      f'131,072 shard + race {shard:.1f} us  65,536 natural {small:.1f} us  loss_sum {out[1][15].item():.12e}', flush=True)
