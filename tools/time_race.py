"""Kernel time of strong-scaling shards (every 16th / 8th / 4th sample of the cost-ordered 1M bench batch) through the training
entry point, with and without racing warps for the expensive head (DPLL_LOSS_RACE).  CUDA events; not the bench."""
import sys, torch
sys.path.insert(0, '.')
import bench
from dair_pll_b200 import ops
dev = torch.device('cuda', 0)
system = bench.make_system(dev, torch.float64)
X, XP = bench.make_batch(system, 1 << 20, 0, dev, torch.float64)
lt, ct = system.multibody_terms.lagrangian_terms, system.multibody_terms.contact_terms
leaves = [t.detach() for t in (lt.inertial_parameters, ct.friction_params, ct.geometries[0].length_params)]
hint = ops.cube_loss_leaf_dp_raw(X, XP, *leaves, bench.DT, 1e-3, want_iters=True)[4]
order = torch.argsort(hint, descending=True, stable=True)
def t_us(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3
for world in (16, 8, 4):
    idx = order[0::world]
    x, xp = X.index_select(0, idx).contiguous(), XP.index_select(0, idx).contiguous()
    row = [t_us(lambda: ops.cube_loss_leaf_dp_raw(x, xp, *leaves, bench.DT, 1e-3, flags=f)) for f in (ops.LOSS_DYNAMIC, ops.LOSS_DYNAMIC | ops.LOSS_RACE)]
    it = ops.cube_loss_leaf_dp_raw(x, xp, *leaves, bench.DT, 1e-3, flags=ops.LOSS_DYNAMIC | ops.LOSS_RACE, want_iters=True)[4]
    print(f'shard 1/{world} ({x.shape[0]} pairs): dynamic {row[0]:.1f} us   dynamic + race {row[1]:.1f} us   max iters in head {int(it[:x.shape[0]//64].max())} rest {int(it[x.shape[0]//64:].max())}', flush=True)
