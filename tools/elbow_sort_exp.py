"""Headroom check for the (one-sample-per-thread) elbow loss kernel: the same batch in input order, sorted by the
true Newton iteration count, and sorted by a crude key (base height).  Measured on B200, B = 262,144:
3.06 / 1.44 / 2.29 ms -- lane divergence costs 2.1x, which wavefront scheduling (as for the cube) would recover."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dair_pll_b200 import ops, synthetic
from dair_pll_b200.inertia import InertialParameterConverter as IPC
from dair_pll_b200.multibody_learnable_system import MultibodyLearnableSystem
dev = torch.device('cuda', 0)
B=262144
s = MultibodyLearnableSystem({'elbow': os.path.join(ROOT, 'dair_pll_b200', 'assets', 'elbow.urdf')}, 0.0068)
pi, fr, half = synthetic.elbow_learnables_perturbed(0)
s.load_state_dict({'multibody_terms.lagrangian_terms.inertial_parameters': IPC.pi_cm_to_theta(pi),
                   'multibody_terms.contact_terms.friction_params': fr,
                   'multibody_terms.contact_terms.geometries.0.length_params': half[0].reshape(1, 3),
                   'multibody_terms.contact_terms.geometries.1.length_params': half[1].reshape(1, 3)})
s = s.to(dev)
x = synthetic.elbow_states(B, seed=0, device=dev)
with torch.no_grad():
    traj, _ = s.simulate(x.unsqueeze(-2), torch.zeros(B, 1, device=dev), 1)
xp = synthetic.perturb_next_state(traj[:, 1], seed=1, n_q=8)
inertia, mu, hl, kin = (t.detach() for t in s._elbow_params(torch.float64, dev))
def timeit(xx, xxp):
    for _ in range(2): ops.elbow_loss_raw(xx, xxp, inertia, mu, hl, kin, 0.0068, 1e-3)
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(5): ops.elbow_loss_raw(xx, xxp, inertia, mu, hl, kin, 0.0068, 1e-3)
    en.record(); torch.cuda.synchronize()
    return st.elapsed_time(en)/5
out = ops.elbow_loss_raw(x, xp, inertia, mu, hl, kin, 0.0068, 1e-3, want_iters=True)
it = out[4]
print('unsorted', timeit(x, xp), 'ms; trivial frac', (it==0).double().mean().item(), 'mean iters nontrivial', it[it>0].double().mean().item())
o = torch.argsort(it)
print('sorted by true iters', timeit(x[o].contiguous(), xp[o].contiguous()), 'ms')
# cheap key: lowest z of the 16 box corners is not available here; use the base height as a crude key
o2 = torch.argsort(xp[:, 6])
print('sorted by base height', timeit(x[o2].contiguous(), xp[o2].contiguous()), 'ms')
