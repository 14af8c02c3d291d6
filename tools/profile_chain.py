"""One warp walking one long Newton chain (a batch of 32 copies of the hardest sample): the command profiled under
ncu to see where a lone warp's cycles go.  Not a benchmark."""
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
idx = int(it.argmax())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x, xp = X[idx:idx + 1].expand(B, 13).contiguous(), XP[idx:idx + 1].expand(B, 13).contiguous()
for _ in range(4):
    out = ops.cube_loss_leaf_dp_raw(x, xp, *leaves, bench.DT, 1e-3, want_iters=True)
torch.cuda.synchronize()
print('iters', int(out[4][0]), 'B', B)
