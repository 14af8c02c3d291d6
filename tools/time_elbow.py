"""Kernel-only timing of the elbow loss kernel (config 3 batch size), CUDA events."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dair_pll_b200 import ops, synthetic  # noqa: E402
from dair_pll_b200.inertia import InertialParameterConverter as IPC  # noqa: E402
from dair_pll_b200.multibody_learnable_system import MultibodyLearnableSystem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=262144)
ap.add_argument('--reps', type=int, default=5)
a = ap.parse_args()
dev = torch.device('cuda', 0)
s = MultibodyLearnableSystem({'elbow': os.path.join(ROOT, 'dair_pll_b200', 'assets', 'elbow.urdf')}, 0.0068)
pi, fr, half = synthetic.elbow_learnables_perturbed(0)
s.load_state_dict({'multibody_terms.lagrangian_terms.inertial_parameters': IPC.pi_cm_to_theta(pi),
                   'multibody_terms.contact_terms.friction_params': fr,
                   'multibody_terms.contact_terms.geometries.0.length_params': half[0].reshape(1, 3),
                   'multibody_terms.contact_terms.geometries.1.length_params': half[1].reshape(1, 3)})
s = s.to(dev)
x = synthetic.elbow_states(a.batch, seed=0, device=dev)
with torch.no_grad():
    traj, _ = s.simulate(x.unsqueeze(-2), torch.zeros(a.batch, 1, device=dev), 1)
xp = synthetic.perturb_next_state(traj[:, 1], seed=1, n_q=8)
inertia, mu, hl, kin = (t.detach() for t in s._elbow_params(torch.float64, dev))
for _ in range(2):
    out = ops.elbow_loss_raw(x, xp, inertia, mu, hl, kin, 0.0068, 1e-3, want_iters=True)
torch.cuda.synchronize()
st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st.record()
for _ in range(a.reps):
    ops.elbow_loss_raw(x, xp, inertia, mu, hl, kin, 0.0068, 1e-3)
en.record()
torch.cuda.synchronize()
ms = st.elapsed_time(en) / a.reps
it = out[4].double()
print(f'elbow loss B={a.batch}: {ms:.3f} ms  {a.batch / ms / 1e3:.2f} M samples/s  mean iters {it.mean().item():.2f} '
      f'max {int(it.max())}  loss_sum {out[2].item():.9e}')
