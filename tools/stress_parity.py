"""One-off robustness sweep: the cube loss kernel against the CPU oracle for strongly perturbed parameters
(mass, inertia, friction, box size scaled by 0.3x .. 3x) and time steps; prints the worst relative deviations."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from dair_pll_b200 import ops, synthetic  # noqa: E402
from oracle import contactnets_oracle as co  # noqa: E402
from oracle.callables import CUBE_TREE, TreeCallables  # noqa: E402

dev = torch.device('cuda', 0)
calls = TreeCallables(CUBE_TREE)
rng = np.random.default_rng(0)
worst = 0.0
for trial in range(8):
    pi, fr, half = synthetic.cube_learnables_perturbed(trial)
    scale_m, scale_i, scale_f, scale_h = (float(np.exp(rng.uniform(np.log(0.3), np.log(3.0)))) for _ in range(4))
    pi = pi.clone()
    pi[..., 0] *= scale_m
    pi[..., 4:7] *= scale_i
    fr = fr * scale_f
    half = half * scale_h
    dt = float(rng.choice([0.002, 0.0068, 0.02]))
    P = co.OracleParams(co.pi_cm_to_theta(pi), fr.clone(), [half.reshape(1, 3).clone()]).requires_grad_()
    n = 20000
    x = synthetic.cube_states(n, seed=100 + trial)
    x[:, 6] *= scale_h                                      # keep the cube near the ground for its new size
    with torch.no_grad():
        xp = synthetic.perturb_next_state(co.sim_step(calls, P, x, dt), seed=200 + trial)
    lo = co.contactnets_loss(calls, P, x, xp, dt)
    lo.sum().backward()
    inertia = co.theta_to_inertia_vector(P.inertial_parameters.detach()).reshape(10).to(dev)
    m = fr.abs()
    mu = (2 * m[0] * m[1] / (m[0] + m[1])).reshape(1).to(dev)
    hl = half.abs().reshape(3).to(dev)
    loss, grad, loss_sum, _, iters = ops.cube_loss_raw(x.to(dev), xp.to(dev), inertia, mu, hl, dt, 1e-3, want_iters=True)
    a, b = loss.cpu().numpy(), lo.detach().numpy()
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-9 * np.abs(b).max())
    # chain rule of the kernel's callable-level gradient to theta for a gradient comparison
    theta = P.inertial_parameters.detach().clone().requires_grad_()
    co.theta_to_inertia_vector(theta).reshape(10).backward(grad[:10].cpu())
    gref = P.inertial_parameters.grad.numpy()
    grel = np.abs(theta.grad.numpy() - gref).max() / np.abs(gref).max()
    worst = max(worst, rel.max(), grel)
    print(f'trial {trial}: m x{scale_m:.2f} I x{scale_i:.2f} mu x{scale_f:.2f} h x{scale_h:.2f} dt {dt}: '
          f'loss rel {rel.max():.2e}  theta-grad rel {grel:.2e}  iters mean {iters.double().mean().item():.2f} max {int(iters.max())} '
          f'masked {(a == 0).mean():.3f}')
print('worst', worst)
