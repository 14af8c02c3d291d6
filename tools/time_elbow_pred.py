"""Prediction-loss path of the two-body system (rollout + backward): boxes (43 tangent directions over the whole rollout) and
learned geometry (61 directions per step through autograd).  CUDA events; not the bench."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dair_pll_b200 import synthetic  # noqa: E402
from dair_pll_b200.multibody_learnable_system import MultibodyLearnableSystem  # noqa: E402

dev = 'cuda:0'
torch.manual_seed(0)
for urdf, cases in (('elbow.urdf', ((4096, 80), (1024, 20))), ('elbow_mesh.urdf', ((4096, 20), (1024, 20)))):
    s = MultibodyLearnableSystem({'elbow': os.path.join(ROOT, 'dair_pll_b200', 'assets', urdf)}, 0.0068).to(dev)
    for n, steps in cases:
        x0 = synthetic.elbow_states(n, seed=41, device=dev).requires_grad_()
        carry = torch.zeros(n, 1, device=dev)

        def step():
            for p in s.parameters():
                p.grad = None
            x0.grad = None
            traj, _ = s.simulate(x0.unsqueeze(-2), carry, steps)
            (traj[:, 1:] ** 2).mean().backward()
        step()
        torch.cuda.synchronize()
        t = time.perf_counter()
        step()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t) * 1e3
        with torch.no_grad():
            s.simulate(x0.unsqueeze(-2), carry, steps)
            torch.cuda.synchronize()
            t = time.perf_counter()
            s.simulate(x0.unsqueeze(-2), carry, steps)
            torch.cuda.synchronize()
            msf = (time.perf_counter() - t) * 1e3
        print(f'{urdf} prediction loss {n} x {steps}: forward {msf:.1f} ms, forward+backward {ms:.1f} ms', flush=True)
