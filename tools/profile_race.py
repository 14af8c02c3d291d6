"""A few launches of the wavefront kernel with racing warps on a 131,072-pair shard of the cost-ordered bench batch (then two
without racing) -- the command profiled under ncu / compute-sanitizer.  Not a benchmark."""
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
idx = order[0::8]
x, xp = X.index_select(0, idx).contiguous(), XP.index_select(0, idx).contiguous()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(3):
    ops.cube_loss_leaf_dp_raw(x, xp, *leaves, bench.DT, 1e-3, flags=ops.LOSS_DYNAMIC | ops.LOSS_RACE)
for _ in range(2):
    ops.cube_loss_leaf_dp_raw(x, xp, *leaves, bench.DT, 1e-3, flags=ops.LOSS_DYNAMIC)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
