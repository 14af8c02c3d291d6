"""How many direction rows of the support-function networks carry a non-zero cotangent in the elbow-mesh step
(bench.py's config-3 workload)?  Rows with d loss / d witness point == 0 contribute nothing to any weight gradient, so
the networks' backward only has to visit the others.  Not a benchmark."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dair_pll_b200 import ops, synthetic  # noqa: E402
from dair_pll_b200.multibody_learnable_system import MultibodyLearnableSystem  # noqa: E402

dev = torch.device('cuda', 0)
DT = 0.0068
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
torch.manual_seed(0)
elbow = MultibodyLearnableSystem({'elbow': os.path.join(ROOT, 'dair_pll_b200', 'assets', 'elbow_mesh.urdf')}, DT).to(dev)
xe = synthetic.elbow_states(B, seed=3, device=dev)
with torch.no_grad():
    te, _ = elbow.simulate(xe.unsqueeze(-2), torch.zeros(B, 1, device=dev), 1)
xpm = synthetic.perturb_next_state(te[:, 1], seed=4, n_q=8)
inertia, mu, half, kin = elbow._elbow_params(torch.float64, dev)
with torch.no_grad():
    pts = elbow._elbow_witness_points(xpm[:, :8])
out = ops.elbow_loss_raw(xe, xpm, inertia.detach(), mu.detach(), None, kin, DT, 1e-3, pts=pts, want_grad_pts=True,
                         want_force=True)
gp = out[5]
act = (gp != 0).any(-1)
print('rows', act.numel(), 'active fraction', act.double().mean().item(), 'per network',
      act[:, :4].double().mean().item(), act[:, 4:].double().mean().item())
print('samples with any active row', act.any(-1).double().mean().item())
f = out[3]
print('samples with non-zero force', (f != 0).any(-1).double().mean().item())
mag = gp.abs().amax(-1)[act]
print('|gp| quantiles of active rows', torch.quantile(mag[:1000000], torch.tensor([0.0, 0.01, 0.5, 0.99, 1.0], device=dev, dtype=mag.dtype)).tolist())
