"""Warm-started solves: kernel time and Newton counts of the cube loss when every sample's solve starts from the optimum
found with slightly different parameters (what a training loop has from its previous epoch).  Not a benchmark."""
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
leaves0 = [t.detach().clone() for t in (lt.inertial_parameters, ct.friction_params, ct.geometries[0].length_params)]


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


for B in (1 << 20, 131072):
    x, xp = X[:B].contiguous(), XP[:B].contiguous()
    cold = ops.cube_loss_leaf_dp_raw(x, xp, *leaves0, bench.DT, 1e-3, want_iters=True, want_u=True)
    order = torch.argsort(cold[4], descending=True, stable=True)
    xo, xpo, u0 = x[order].contiguous(), xp[order].contiguous(), cold[5][order].contiguous()
    reps = 50 if B > 500000 else 200
    ms_cold = t_ms(lambda: ops.cube_loss_leaf_dp_raw(xo, xpo, *leaves0, bench.DT, 1e-3, flags=ops.LOSS_DYNAMIC), reps)
    print(f'B={B}: cold start (cost order, dynamic) {ms_cold:.4f} ms, mean iters {cold[4].double().mean().item():.3f}', flush=True)
    for rel in (0.0, 1e-4, 1e-3, 1e-2):
        g = torch.Generator(device='cpu').manual_seed(1)
        leaves = [t * (1 + rel * (2 * torch.rand(t.shape, generator=g, dtype=torch.float64).to(dev) - 1)) for t in leaves0]
        ref = ops.cube_loss_leaf_dp_raw(xo, xpo, *leaves, bench.DT, 1e-3, flags=ops.LOSS_DYNAMIC, want_iters=True)
        warm = ops.cube_loss_leaf_dp_raw(xo, xpo, *leaves, bench.DT, 1e-3, flags=ops.LOSS_DYNAMIC, want_iters=True, u_init=u0,
                                         want_u=True)
        # re-order by the warm-started cost (the hint a training loop would carry)
        o2 = torch.argsort(warm[4], descending=True, stable=True)
        xw, xpw, uw = xo[o2].contiguous(), xpo[o2].contiguous(), u0[o2].contiguous()
        ms_w = t_ms(lambda: ops.cube_loss_leaf_dp_raw(xw, xpw, *leaves, bench.DT, 1e-3, flags=ops.LOSS_DYNAMIC, u_init=uw,
                                                      want_u=True), reps)
        dl = (warm[0] - ref[0]).abs().max().item()
        dg = ((warm[1][:15] - ref[1][:15]).abs().max() / ref[1][:15].abs().max()).item()
        print(f'   parameters moved by {rel:g}: warm start {ms_w:.4f} ms, mean iters {warm[4].double().mean().item():.3f} '
              f'(max {int(warm[4].max())}), max |dloss| {dl:.1e}, grad rel {dg:.1e}', flush=True)
